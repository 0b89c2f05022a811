import numpy as _np
from scipy.spatial.transform import Rotation as _R

from ..._core import asarray as _asarray


class Rotation:
    def __init__(self, r):
        self._r = r

    @classmethod
    def from_euler(cls, seq, angles, degrees=False):
        return cls(_R.from_euler(seq, _np.asarray(angles), degrees=degrees))

    def as_quat(self, **kw):
        return _asarray(self._r.as_quat(**kw))  # xyzw, like jax.scipy

    def as_matrix(self):
        return _asarray(self._r.as_matrix())
