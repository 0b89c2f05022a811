"""jax.random over NumPy's Philox (NOT JAX's threefry stream: goldens carry their inputs explicitly)."""
import numpy as _np

from ._core import asarray as _asarray


def PRNGKey(seed):
    return _asarray(_np.array([0, int(seed)], dtype=_np.uint32))


key = PRNGKey


def split(key, num=2):
    k = _np.asarray(key)
    rng = _np.random.Generator(_np.random.Philox(key=int(k[0]) * (1 << 32) + int(k[1])))
    return _asarray(rng.integers(0, 2**32, size=(num, 2), dtype=_np.uint32))


def uniform(key, shape=(), dtype=float, minval=0.0, maxval=1.0):
    k = _np.asarray(key)
    rng = _np.random.Generator(_np.random.Philox(key=int(k[0]) * (1 << 32) + int(k[1])))
    lo, hi = _np.asarray(minval, dtype=_np.float64), _np.asarray(maxval, dtype=_np.float64)
    return _asarray(lo + (hi - lo) * rng.random(size=tuple(shape)))
