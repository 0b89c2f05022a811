"""`jax.numpy` over NumPy: every function returns `Array` views (see _core.py)."""
import sys as _sys
import types as _types

import numpy as _np

from ._core import Array, asarray as _asarray, wrap_fn as _wrap_fn, wrap_out as _wrap_out

ndarray = Array
newaxis = None
pi = _np.pi
inf = _np.inf
nan = _np.nan
e = _np.e
s_ = _np.s_
index_exp = _np.index_exp
float32, float64, int32, int64, bool_, uint8, uint32, int8, int16, float16 = (
    _np.float32, _np.float64, _np.int32, _np.int64, _np.bool_, _np.uint8, _np.uint32, _np.int8, _np.int16, _np.float16)
finfo, iinfo, dtype, result_type, issubdtype, floating, integer = (
    _np.finfo, _np.iinfo, _np.dtype, _np.result_type, _np.issubdtype, _np.floating, _np.integer)


def asarray(x, dtype=None, **_):
    return _asarray(x, dtype)


def array(x, dtype=None, copy=True, **_):
    return _asarray(_np.array(x, dtype=(_np.float64 if dtype is float else _np.int64 if dtype is int else dtype), copy=True))


def astype(x, dtype, **_):
    return _asarray(x).astype(dtype)


class _Linalg(_types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _wrap_fn(getattr(_np.linalg, name))

    @staticmethod
    def lstsq(a, b, rcond=None, **_):
        return _wrap_out(tuple(_np.linalg.lstsq(_np.asarray(a), _np.asarray(b), rcond=rcond)))


linalg = _Linalg("jax.numpy.linalg")
_sys.modules["jax.numpy.linalg"] = linalg


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    fn = getattr(_np, name)
    if callable(fn) and not isinstance(fn, type):
        w = _wrap_fn(fn)
        globals()[name] = w
        return w
    return fn
