"""Functional API mirroring ``jaxsim.api`` for the hot path (``model``, ``data``, ``common``)."""
from . import common, contact, data, frame, integrators, kin_dyn_parameters, model, ode  # noqa: F401
