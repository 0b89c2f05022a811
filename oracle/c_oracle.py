"""ctypes wrapper of the plain-C oracle (``oracle/c/jaxsim_oracle.c``).  TEST INFRASTRUCTURE
ONLY -- same rules as ``oracle/jaxsim_oracle.py``.  float64, SoftContacts / no contacts,
SemiImplicitEuler."""

from __future__ import annotations

import ctypes as C
import os
import pathlib
import subprocess

import numpy as np

from jaxsim_b200 import _lib as _abi  # only for the ctypes mirror of struct B200SimModelDesc

from . import jaxsim_oracle as O

_HERE = pathlib.Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liboracle.so"
_lib = None


def load():
    global _lib
    if _lib is None:
        if not _SO.exists():
            subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
        _lib = C.CDLL(str(_SO))
        _lib.oracle_step.restype = C.c_int
        _lib.oracle_step.argtypes = [C.POINTER(_abi.B200SimModelDesc), C.c_int64, C.c_int] + [C.c_void_p] * 20
    return _lib


def _desc(om: O.OracleModel):
    kd = om.kin_dyn_parameters
    nL, n = kd.number_of_links(), kd.number_of_joints()
    cp = kd.contact_parameters
    nc = len(cp.body)
    keep = {}

    def dp(name, arr):
        keep[name] = np.ascontiguousarray(arr, dtype=np.float64)
        return keep[name].ctypes.data_as(_abi.c_dp)

    def ip(name, arr):
        keep[name] = np.ascontiguousarray(arr, dtype=np.int32)
        return keep[name].ctypes.data_as(_abi.c_ip)

    axis = np.zeros((nL, 3))
    if n:
        axis[1:] = kd.joint_model.joint_axis
    jp = kd.joint_parameters
    big = np.finfo(np.float64).max
    d = _abi.B200SimModelDesc(
        abi_version=1, n_links=nL, n_dofs=n, n_points=nc, floating_base=int(om.floating_base),
        contact_model=1 if (om.contact_model == "soft" and nc > 0) else 0, enable_friction=int(om.enable_friction), reserved0=0,
        parent=ip("parent", kd.parent_array), joint_type=ip("jt", kd.joint_model.joint_types),
        lam_H_pre=dp("lam", kd.joint_model.lam_H_pre), suc_H_i=dp("suc", kd.joint_model.suc_H_i), joint_axis=dp("axis", axis),
        link_mass=dp("mass", kd.link_parameters.mass), link_com=dp("com", kd.link_parameters.center_of_mass),
        link_inertia=dp("inertia", kd.link_parameters.inertia_elements),
        friction_static=dp("kc", jp.friction_static), friction_viscous=dp("kv", jp.friction_viscous),
        position_limits_min=dp("smin", np.clip(jp.position_limits_min, -big, big)),
        position_limits_max=dp("smax", np.clip(jp.position_limits_max, -big, big)),
        position_limit_spring=dp("ks", jp.position_limit_spring), position_limit_damper=dp("kd", jp.position_limit_damper),
        point_body=ip("pb", np.array(cp.body, dtype=np.int32)),
        point_position=dp("pp", np.asarray(cp.point, dtype=float).reshape(-1, 3) if nc else np.zeros((0, 3))),
        point_enabled=ip("pe", np.array(cp.enabled, dtype=np.int32)),
        time_step=om.time_step, gravity=om.gravity, terrain_height=om.terrain_height,
        soft_K=om.K, soft_D=om.D, soft_mu=om.mu, soft_p=om.p, soft_q=om.q,
        torque_max=om.torque_max, omega_th=om.omega_th, omega_max=om.omega_max,
    )
    return d, keep


def step(om: O.OracleModel, data: O.OracleData, link_forces_inertial=None, joint_force_references=None,
         nthreads: int | None = None, caches: bool = True) -> O.OracleData:
    """Same contract as ``oracle.jaxsim_oracle.step`` (float64), computed by the C oracle."""
    lib = load()
    d, keep = _desc(om)
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    s, sd, q = f(data.joint_positions), f(data.joint_velocities), f(data.base_quaternion)
    vl, w, p = f(data.base_linear_velocity), f(data.base_angular_velocity), f(data.base_position)
    m, tau, fx = f(data.tangential_deformation), f(joint_force_references), f(link_forces_inertial)
    B, nL, n, nc = q.shape[0], om.number_of_links(), om.dofs(), len(om.kin_dyn_parameters.contact_parameters.body)
    E = np.empty
    s_o, sd_o, q_o, vl_o, w_o, p_o = E((B, n)), E((B, n)), E((B, 4)), E((B, 3)), E((B, 3)), E((B, 3))
    m_o = E((B, nc, 3))
    W_H_B, iXl, W_H_L, W_v = (E((B, 4, 4)), E((B, nL, 6, 6)), E((B, nL, 4, 4)), E((B, nL, 6))) if caches else (None,) * 4
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
    nt = nthreads if nthreads is not None else (os.cpu_count() or 1)
    rc = lib.oracle_step(C.byref(d), B, int(nt), ptr(s), ptr(sd), ptr(q), ptr(vl), ptr(w), ptr(p), ptr(m), ptr(tau), ptr(fx),
                         ptr(s_o), ptr(sd_o), ptr(q_o), ptr(vl_o), ptr(w_o), ptr(p_o), ptr(m_o),
                         ptr(W_H_B), ptr(iXl), ptr(W_H_L), ptr(W_v))
    if rc != 0:
        raise RuntimeError(f"oracle_step -> {rc}")
    return O.OracleData(
        joint_positions=s_o, joint_velocities=sd_o, base_quaternion=q_o, base_linear_velocity=vl_o,
        base_angular_velocity=w_o, base_position=p_o, base_transform=W_H_B, joint_transforms=iXl,
        link_transforms=W_H_L, link_velocities=W_v, tangential_deformation=m_o,
    )
