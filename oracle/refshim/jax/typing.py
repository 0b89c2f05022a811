import typing as _t

import numpy as _np

from ._core import Array

ArrayLike = _t.Union[Array, _np.ndarray, _np.generic, bool, int, float, complex]
DTypeLike = _t.Any
