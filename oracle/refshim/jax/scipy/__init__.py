from . import linalg, spatial  # noqa: F401
