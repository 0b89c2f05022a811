"""Functional API mirroring ``jaxsim.api`` for the hot path (``model``, ``data``, ``common``)."""
from . import common, data, kin_dyn_parameters, model  # noqa: F401
