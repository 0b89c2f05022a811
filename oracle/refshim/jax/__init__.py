"""NumPy-backed stand-in for `jax` (test infrastructure only, see ../README.md)."""
import contextlib as _contextlib
import types as _types

from . import _core
from ._core import Array, custom_jvp, jit, vmap
from . import numpy, lax, tree_util, tree, flatten_util, random, typing, scipy, core, interpreters, debug, _src  # noqa: F401

__version__ = "0.0.0-numpy-standin"


class _Config:
    jax_enable_x64 = True

    def update(self, key, value):
        setattr(self, key, value)


config = _Config()


class _Device:
    platform = "cpu"
    device_kind = "cpu"
    id = 0

    def __repr__(self):
        return "CpuDevice(id=0, numpy stand-in)"


def devices(*_a, **_k):
    return [_Device()]


def default_backend():
    return "cpu"


def named_scope(_name):
    return _contextlib.nullcontext()


def block_until_ready(x):
    return x


def device_put(x, *_a, **_k):
    return x


def device_get(x):
    return x


def grad(*_a, **_k):
    raise NotImplementedError("the NumPy stand-in executes forward computations only")


jacfwd = jacrev = jvp = vjp = value_and_grad = grad
