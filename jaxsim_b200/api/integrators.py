"""Integrators beyond the fused semi-implicit Euler step.

``rk4_integration`` (``src/jaxsim/api/integrators.py:91-156``) is one call of ``b200sim_step_rk4``: the four
``system_dynamics`` evaluations, the stage updates and the cache refresh of ``data.replace`` run as launches on the
current stream with the stage state in device scratch.  ``rk4fast_integration`` (``:159-263``) freezes the contact forces
and glues four ``b200sim_aba`` launches with elementwise torch ops.  The semi-implicit Euler scheme (``:14-88``) lives
entirely inside ``b200sim_step``.
"""

from __future__ import annotations

import torch

from jaxsim_b200 import _lib
from jaxsim_b200.rbda.contacts import SoftContacts

from . import ode
from .common import VelRepr, other_representation_to_inertial


def _tree(f, *trees):
    out = {}
    for k in trees[0]:
        if isinstance(trees[0][k], dict):
            out[k] = _tree(f, *[t[k] for t in trees])
        else:
            out[k] = f(*[t[k] for t in trees])
    return out


def rk4_integration(model, data, link_forces_inertial, joint_force_references):
    """``step`` with ``IntegratorType.RungeKutta4`` (``api/integrators.py:91-156``) for batched data: one call of
    ``b200sim_step_rk4`` -- the actuation model, the four ``system_dynamics`` evaluations, the stage updates and the
    ``data.replace`` of the result (normalised quaternion, caches) run as ten launches on the current stream with the
    stage state in device scratch; nothing returns to the host in between."""
    from .model import _alloc_outputs, _dtype_code, _ptr, _stream_ptr

    q = data._base_quaternion.contiguous()
    dev, dtype = q.device, q.dtype
    dm = model.device_model(dev)
    B = q.shape[0]
    nL, n, nc = model.number_of_links(), model.dofs(), model.number_of_collidable_points()
    soft = isinstance(model.contact_model, SoftContacts)
    c = lambda t: None if t is None else torch.as_tensor(t, dtype=dtype, device=dev).contiguous()  # noqa: E731
    m = c(data.contact_state.get("tangential_deformation")) if (soft and data.contact_state) else None
    tau, fext = c(joint_force_references), c(link_forces_inertial)
    if tau is not None and tau.shape != (B, n):
        raise ValueError(tau.shape, (B, n))
    if fext is not None and fext.shape != (B, nL, 6):
        raise ValueError(fext.shape, (B, nL, 6))
    o = _alloc_outputs(model, B, dtype, dev, True, soft)
    g = o.get
    with torch.cuda.device(dev):
        rc = _lib.load().b200sim_step_rk4(
            dm.handle, _dtype_code(dtype), B, _ptr(c(data._joint_positions)), _ptr(c(data._joint_velocities)), _ptr(q),
            _ptr(c(data._base_linear_velocity)), _ptr(c(data._base_angular_velocity)), _ptr(c(data._base_position)), _ptr(m),
            _ptr(tau), _ptr(fext), _ptr(o["s"]), _ptr(o["sd"]), _ptr(o["q"]), _ptr(o["vl"]), _ptr(o["om"]), _ptr(o["p"]),
            _ptr(g("m")), _ptr(g("W_H_B")), _ptr(g("iXl")), _ptr(g("W_H_L")), _ptr(g("W_v")), _stream_ptr(dev),
        )
    _lib.check(rc, "b200sim_step_rk4")
    contact_state = dict(data.contact_state) if data.contact_state else {}
    if soft:
        contact_state["tangential_deformation"] = o["m"]
    from .data import JaxSimModelData

    return JaxSimModelData(
        velocity_representation=data.velocity_representation,
        _joint_positions=o["s"], _joint_velocities=o["sd"], _base_quaternion=o["q"], _base_linear_velocity=o["vl"],
        _base_angular_velocity=o["om"], _base_position=o["p"], _base_transform=g("W_H_B"), _joint_transforms=g("iXl"),
        _link_transforms=g("W_H_L"), _link_velocities=g("W_v"), contact_state=contact_state,
    )


def rk4fast_integration(model, data, link_forces_inertial, joint_torques):
    """``rk4fast_integration`` (``api/integrators.py:159-263``) for batched data, quirks included: the contact forces
    are evaluated once at the initial state and held over the four stages; the integrated contact state starts from
    ``update_contact_state(m_dot)`` = the derivative itself (``soft.py:164-176``) and each stage returns its contact
    state as the "derivative"; the position derivatives come from the ORIGINAL data (``:206-209``).  Each stage is one
    ABA launch (``b200sim_aba``) on the stage's state with the frozen link forces."""
    from . import contact as _contact
    from . import model as _model
    from .data import JaxSimModelData
    from jaxsim_b200.rbda.contacts import SoftContacts

    if not isinstance(model.contact_model, SoftContacts) or model.number_of_collidable_points() == 0:
        raise NotImplementedError(
            "RungeKutta4Fast: the reference only runs with collidable points (api/integrators.py:175-190); "
            "implemented for SoftContacts")
    dt = model.time_step
    q = data._base_quaternion
    nrm = torch.linalg.norm(q, dim=-1, keepdim=True)
    q = q / torch.where(nrm == 0, torch.ones_like(nrm), nrm)
    W_f_L, m_dot = _contact.soft_link_contact_forces(model, data)
    W_f_total = W_f_L if link_forces_inertial is None else W_f_L + link_forces_inertial
    # system_position_dynamics(data) with Baumgarte K = 1.0 (api/ode.py:134-171), of the ORIGINAL data
    w, p = data._base_angular_velocity, data._base_position
    pd = data._base_linear_velocity + torch.linalg.cross(w, p)
    qn = data.base_orientation
    nw = torch.linalg.norm(w, dim=-1, keepdim=True)
    nq = torch.linalg.norm(qn, dim=-1, keepdim=True)
    v0 = nw * (1.0 - nq)
    qw, qx, qy, qz = qn.unbind(-1)
    wx, wy, wz = w.unbind(-1)
    v0 = v0.squeeze(-1)
    qd = 0.5 * torch.stack([qw * v0 - qx * wx - qy * wy - qz * wz, qx * v0 + qw * wx + qz * wy - qy * wz,
                            qy * v0 - qz * wx + qw * wy + qx * wz, qz * v0 + qy * wx - qx * wy + qw * wz], dim=-1)
    sd0 = data._joint_velocities

    def f(x):
        d = JaxSimModelData(
            velocity_representation=VelRepr.Inertial, _joint_positions=x["joint_positions"], _joint_velocities=x["joint_velocities"],
            _base_quaternion=x["base_quaternion"], _base_linear_velocity=x["base_linear_velocity"],
            _base_angular_velocity=x["base_angular_velocity"], _base_position=x["base_position"], contact_state={})
        vd, sdd = _model.forward_dynamics_aba(model, d, joint_forces=joint_torques, link_forces=W_f_total)
        return dict(base_position=pd, base_quaternion=qd, joint_positions=sd0, base_linear_velocity=vd[:, 0:3],
                    base_angular_velocity=vd[:, 3:6], joint_velocities=sdd,
                    contact_state={"tangential_deformation": x["contact_state"]["tangential_deformation"]})

    x0 = dict(
        base_position=data._base_position, base_quaternion=q, joint_positions=data._joint_positions,
        base_linear_velocity=data._base_linear_velocity, base_angular_velocity=data._base_angular_velocity,
        joint_velocities=data._joint_velocities, contact_state={"tangential_deformation": m_dot},
    )
    mid = lambda x, d: x + (0.5 * dt) * d  # noqa: E731
    fin = lambda x, d: x + dt * d  # noqa: E731
    k1 = f(x0)
    k2 = f(_tree(mid, x0, k1))
    k3 = f(_tree(mid, x0, k2))
    k4 = f(_tree(fin, x0, k3))
    dxdt = _tree(lambda a, b, c, d: (a + 2 * b + 2 * c + d) / 6, k1, k2, k3, k4)
    xf = _tree(fin, x0, dxdt)
    return data.replace(
        model, joint_positions=xf["joint_positions"], joint_velocities=xf["joint_velocities"],
        base_quaternion=xf["base_quaternion"], base_position=xf["base_position"],
        contact_state=xf["contact_state"],
        _inertial_base_velocity=(xf["base_linear_velocity"], xf["base_angular_velocity"]),
    )


def step_rk4(model, data, *, link_forces=None, joint_force_references=None):
    """``js.model.step`` with ``IntegratorType.RungeKutta4`` / ``RungeKutta4Fast`` (``api/model.py:2601-2681``)."""
    if data._base_quaternion.dim() == 1:  # unbatched data, like every other entry point
        from .data import _map_leaves

        one = _map_leaves(data, lambda t: t.unsqueeze(0))
        lf = None if link_forces is None else torch.as_tensor(link_forces, dtype=data._base_quaternion.dtype, device=data._base_quaternion.device).unsqueeze(0)
        jf = None if joint_force_references is None else torch.as_tensor(joint_force_references, dtype=data._base_quaternion.dtype, device=data._base_quaternion.device).unsqueeze(0)
        return _map_leaves(step_rk4(model, one, link_forces=lf, joint_force_references=jf), lambda t: t.squeeze(0))
    fext = None
    if link_forces is not None:
        fext = other_representation_to_inertial(
            torch.as_tensor(link_forces, dtype=data._base_quaternion.dtype, device=data._base_quaternion.device),
            data.velocity_representation, data.link_transforms, is_force=True,
        )
    from .model import IntegratorType

    if model.integrator == IntegratorType.RungeKutta4Fast:
        tau = ode.compute_resultant_torques(model, data, joint_force_references=joint_force_references)
        return rk4fast_integration(model, data, fext, tau)
    return rk4_integration(model, data, fext, joint_force_references)
