"""ctypes binding of ``libb200sim.so`` (C ABI declared in ``include/b200sim.h``).

The product path has NO fallback: if the CUDA library is missing or fails to load, every
entry point raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``.
"""

from __future__ import annotations

import ctypes as C
import pathlib

_HERE = pathlib.Path(__file__).resolve().parent
import os as _os

# B200SIM_LIB: alternative build of the same library (A/B timing of compile-time variants)
LIB_PATH = pathlib.Path(_os.environ["B200SIM_LIB"]) if _os.environ.get("B200SIM_LIB") else _HERE / "csrc" / "libb200sim.so"

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class B200SimModelDesc(C.Structure):
    """Mirror of ``struct B200SimModelDesc`` (include/b200sim.h)."""

    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_links", C.c_int32),
        ("n_dofs", C.c_int32),
        ("n_points", C.c_int32),
        ("floating_base", C.c_int32),
        ("contact_model", C.c_int32),
        ("enable_friction", C.c_int32),
        ("reserved0", C.c_int32),
        ("parent", c_ip),
        ("joint_type", c_ip),
        ("lam_H_pre", c_dp),
        ("suc_H_i", c_dp),
        ("joint_axis", c_dp),
        ("link_mass", c_dp),
        ("link_com", c_dp),
        ("link_inertia", c_dp),
        ("friction_static", c_dp),
        ("friction_viscous", c_dp),
        ("position_limits_min", c_dp),
        ("position_limits_max", c_dp),
        ("position_limit_spring", c_dp),
        ("position_limit_damper", c_dp),
        ("point_body", c_ip),
        ("point_position", c_dp),
        ("point_enabled", c_ip),
        ("time_step", C.c_double),
        ("gravity", C.c_double),
        ("terrain_height", C.c_double),
        ("soft_K", C.c_double),
        ("soft_D", C.c_double),
        ("soft_mu", C.c_double),
        ("soft_p", C.c_double),
        ("soft_q", C.c_double),
        ("torque_max", C.c_double),
        ("omega_th", C.c_double),
        ("omega_max", C.c_double),
        ("rigid_regularization", C.c_double),
        ("relaxed_time_constant", C.c_double), ("relaxed_damping_coefficient", C.c_double), ("relaxed_d_min", C.c_double),
        ("relaxed_d_max", C.c_double), ("relaxed_width", C.c_double), ("relaxed_midpoint", C.c_double),
        ("relaxed_power", C.c_double),
    ]


ABI_VERSION = 3
OPT_TMA_STORE = 1
OPT_RIGID_QP_F32 = 2
OPT_GENERIC_KERNEL = 4
OPT_BULK_IN = 8
OPT_NO_PDL = 16
OPT_STEP_V1 = 32
OPT_NO_BULK_IN = 64
OPT_RIGID_MONO = 128
REPR_INERTIAL, REPR_BODY, REPR_MIXED = 0, 1, 2
STATUS_QUATERNION_NAN = 1
STATUS_QUATERNION_NOT_UNIT = 2
STATUS_NON_FINITE = 4
STATUS_QP_NOT_CONVERGED = 8
EXPORTED_SYMBOLS = (
    "b200sim_version",
    "b200sim_model_create",
    "b200sim_model_destroy",
    "b200sim_model_update_link_params",
    "b200sim_model_set_tuning",
    "b200sim_model_set_options",
    "b200sim_model_query",
    "b200sim_step",
    "b200sim_step_n",
    "b200sim_step_n_status",
    "b200sim_step_n_ex",
    "b200sim_fk",
    "b200sim_aba",
    "b200sim_rnea",
    "b200sim_crba",
    "b200sim_step_jvp",
    "b200sim_dynamics",
    "b200sim_step_rk4",
    "b200sim_step_vjp",
    "b200sim_step_jvp_ex",
)

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises ``RuntimeError`` if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    vp = C.c_void_p
    lib.b200sim_version.restype = C.c_char_p
    lib.b200sim_model_create.argtypes = [C.POINTER(B200SimModelDesc), C.c_int, C.POINTER(vp)]
    lib.b200sim_model_create.restype = C.c_int
    lib.b200sim_model_destroy.argtypes = [vp]
    lib.b200sim_model_destroy.restype = None
    lib.b200sim_model_update_link_params.argtypes = [vp, c_dp, c_dp, c_dp]
    lib.b200sim_model_update_link_params.restype = C.c_int
    lib.b200sim_model_set_tuning.argtypes = [vp, C.c_int, C.c_int]
    lib.b200sim_model_set_tuning.restype = C.c_int
    lib.b200sim_model_query.argtypes = [vp, C.c_int, C.c_int64, c_ip, c_ip, c_ip, c_ip]
    lib.b200sim_model_query.restype = C.c_int
    lib.b200sim_model_set_options.argtypes = [vp, C.c_int32]
    lib.b200sim_model_set_options.restype = C.c_int
    lib.b200sim_step_n.argtypes = (
        [vp, C.c_int, C.c_int64, C.c_int32] + [vp] * 8 + [C.c_int64, vp, C.c_int64] + [vp] * 14
    )
    lib.b200sim_step_n.restype = C.c_int
    lib.b200sim_step_n_status.argtypes = (
        [vp, C.c_int, C.c_int64, C.c_int32] + [vp] * 8 + [C.c_int64, vp, C.c_int64] + [vp] * 15
    )
    lib.b200sim_step_n_status.restype = C.c_int
    lib.b200sim_step_n_ex.argtypes = (
        [vp, C.c_int, C.c_int64, C.c_int32] + [vp] * 8 + [C.c_int64, vp, C.c_int64] + [vp] * 13 + [C.c_int32, vp, vp]
    )
    lib.b200sim_step_n_ex.restype = C.c_int
    lib.b200sim_step.argtypes = [vp, C.c_int, C.c_int64] + [vp] * 21
    lib.b200sim_step.restype = C.c_int
    lib.b200sim_fk.argtypes = [vp, C.c_int, C.c_int64] + [vp] * 12
    lib.b200sim_fk.restype = C.c_int
    lib.b200sim_aba.argtypes = [vp, C.c_int, C.c_int64] + [vp] * 11
    lib.b200sim_aba.restype = C.c_int
    lib.b200sim_rnea.argtypes = [vp, C.c_int, C.c_int64] + [vp] * 12
    lib.b200sim_rnea.restype = C.c_int
    lib.b200sim_crba.argtypes = [vp, C.c_int, C.c_int64, vp, vp, vp]
    lib.b200sim_crba.restype = C.c_int
    lib.b200sim_step_jvp.argtypes = [vp, C.c_int64, C.c_int32, c_dp] + [vp] * 20
    lib.b200sim_step_jvp.restype = C.c_int
    lib.b200sim_dynamics.argtypes = [vp, C.c_int, C.c_int64] + [vp] * 15
    lib.b200sim_dynamics.restype = C.c_int
    lib.b200sim_step_rk4.argtypes = [vp, C.c_int, C.c_int64] + [vp] * 21
    lib.b200sim_step_rk4.restype = C.c_int
    lib.b200sim_step_jvp_ex.argtypes = [vp, C.c_int64, C.c_int32, c_dp, C.c_int64, C.c_int32] + [vp] * 20
    lib.b200sim_step_jvp_ex.restype = C.c_int
    lib.b200sim_step_vjp.argtypes = [vp, C.c_int64] + [vp] * 18
    lib.b200sim_step_vjp.restype = C.c_int
    _lib = lib
    return lib


class B200SimError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    names = {-1: "invalid argument", -2: "unsupported by this build", -3: "model too large for shared memory"}
    msg = names.get(rc, f"CUDA error {rc}" if rc > 0 else f"error {rc}")
    raise B200SimError(f"{what} failed: {msg}")
