"""``system_dynamics`` (``src/jaxsim/api/ode.py:174-225``) on the device, and the actuation
model as a host-side torch shim for the integrators that need the resultant torques ahead
of the dynamics call (``api/model.py:2658``, ``api/actuation_model.py:7-126``)."""

from __future__ import annotations

import torch

from jaxsim_b200 import _lib
from jaxsim_b200.rbda.contacts import SoftContacts


def compute_resultant_torques(model, data, *, joint_force_references: torch.Tensor | None = None) -> torch.Tensor:
    """``js.actuation_model.compute_resultant_torques`` (``api/actuation_model.py:7-98``) +
    ``tn_curve_fn`` (``:101-126``) as elementwise torch ops (used by the RK4 path only: the
    semi-implicit step evaluates the same model inside the kernel)."""
    s, sd = data._joint_positions, data._joint_velocities
    dt, dev = s.dtype, s.device
    jp = model.kin_dyn_parameters.joint_parameters
    t = lambda a: torch.as_tensor(a, dtype=dt, device=dev)  # noqa: E731
    tau_ref = torch.zeros_like(s) if joint_force_references is None else torch.as_tensor(joint_force_references, dtype=dt, device=dev)
    lower = torch.clamp(s - t(jp.position_limits_min), max=0.0)
    upper = torch.clamp(s - t(jp.position_limits_max), min=0.0)
    tau_lim = -t(jp.position_limit_spring) * (lower + upper)
    tau_lim = tau_lim - tau_lim * t(jp.position_limit_damper) * sd
    tau_fr = torch.zeros_like(s)
    if model.actuation_params.enable_friction:
        tau_fr = -(t(jp.friction_static) * torch.sign(sd) + t(jp.friction_viscous) * sd)
    ap = model.actuation_params
    av = sd.abs()
    lim = torch.where(
        av <= ap.omega_th, torch.full_like(av, ap.torque_max),
        torch.where(av <= ap.omega_max, ap.torque_max * (1 - (av - ap.omega_th) / (ap.omega_max - ap.omega_th)), torch.zeros_like(av)),
    )
    return torch.minimum(torch.maximum(tau_ref + tau_fr + tau_lim, -lim), lim)


def system_dynamics(model, state: dict, *, link_forces_inertial: torch.Tensor | None = None,
                    joint_torques: torch.Tensor | None = None) -> dict:
    """``js.ode.system_dynamics`` for a batched state dict with the reference's keys
    (``base_position, base_quaternion, joint_positions, base_linear_velocity,
    base_angular_velocity, joint_velocities, contact_state``), inertial-fixed
    representation.  Returns the derivative dict with the same keys
    (``api/ode.py:216-225``)."""
    from .model import _dtype_code, _ptr, _stream_ptr

    q = state["base_quaternion"].contiguous()
    dev, dtype = q.device, q.dtype
    dm = model.device_model(dev)
    B = q.shape[0]
    nL, n, nc = model.number_of_links(), model.dofs(), model.number_of_collidable_points()
    s = state["joint_positions"].contiguous()
    sd = state["joint_velocities"].contiguous()
    vl = state["base_linear_velocity"].contiguous()
    om = state["base_angular_velocity"].contiguous()
    p = state["base_position"].contiguous()
    soft = isinstance(model.contact_model, SoftContacts)
    m = state.get("contact_state", {}).get("tangential_deformation") if soft else None
    m = None if m is None else m.contiguous()
    tau = None if joint_torques is None else joint_torques.contiguous()
    fext = None if link_forces_inertial is None else link_forces_inertial.contiguous()
    new = lambda *shape: torch.empty(shape, dtype=dtype, device=dev)  # noqa: E731
    pd, qd, vd, sdd = new(B, 3), new(B, 4), new(B, 6), new(B, n)
    md = new(B, nc, 3) if soft else None
    with torch.cuda.device(dev):
        rc = _lib.load().b200sim_dynamics(
            dm.handle, _dtype_code(dtype), B, _ptr(s), _ptr(sd), _ptr(q), _ptr(vl), _ptr(om), _ptr(p), _ptr(m),
            _ptr(tau), _ptr(fext), _ptr(pd), _ptr(qd), _ptr(vd), _ptr(sdd), _ptr(md), _stream_ptr(dev),
        )
    _lib.check(rc, "b200sim_dynamics")
    return dict(
        base_position=pd, base_quaternion=qd, joint_positions=sd, base_linear_velocity=vd[:, 0:3],
        base_angular_velocity=vd[:, 3:6], joint_velocities=sdd,
        contact_state={"tangential_deformation": md} if soft else {},
    )
