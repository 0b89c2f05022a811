from ._core import ravel_pytree  # noqa: F401
