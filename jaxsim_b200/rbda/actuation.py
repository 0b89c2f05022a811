"""``ActuationParams`` (``src/jaxsim/rbda/actuation/common.py:10-19``)."""

from __future__ import annotations

import dataclasses


@dataclasses.dataclass(frozen=True)
class ActuationParams:
    torque_max: float = 3000.0  # (Nm)
    omega_th: float = 30.0  # (rad/s)
    omega_max: float = 100.0  # (rad/s)
    enable_friction: bool = True
