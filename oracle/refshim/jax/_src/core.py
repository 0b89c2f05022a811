from ..core import Tracer  # noqa: F401
