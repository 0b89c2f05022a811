import scipy.linalg as _sl

from .._core import wrap_fn as _wrap_fn

block_diag = _wrap_fn(_sl.block_diag)
solve = _wrap_fn(_sl.solve)
cho_factor = _wrap_fn(_sl.cho_factor)
cho_solve = _wrap_fn(_sl.cho_solve)
solve_triangular = _wrap_fn(_sl.solve_triangular)
cholesky = _wrap_fn(_sl.cholesky)
