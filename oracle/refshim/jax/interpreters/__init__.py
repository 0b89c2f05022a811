from . import partial_eval  # noqa: F401
