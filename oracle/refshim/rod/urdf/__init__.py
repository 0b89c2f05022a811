from . import exporter  # noqa: F401
