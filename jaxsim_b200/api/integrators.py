"""Integrators beyond the fused semi-implicit Euler step.

``rk4_integration`` restates ``src/jaxsim/api/integrators.py:91-156``: four evaluations of
``system_dynamics`` (one kernel launch each, ``b200sim_dynamics``) combined with elementwise
torch ops, then the cache refresh of ``data.replace`` (``b200sim_fk``).  The semi-implicit
Euler scheme (``:14-88``) lives entirely inside ``b200sim_step``.
"""

from __future__ import annotations

import torch

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


def rk4_integration(model, data, link_forces_inertial, joint_torques):
    """``rk4_integration`` (``api/integrators.py:91-156``) for batched data."""
    dt = model.time_step
    q = data._base_quaternion
    nrm = torch.linalg.norm(q, dim=-1, keepdim=True)
    q = q / torch.where(nrm == 0, torch.ones_like(nrm), nrm)

    def f(x):
        # data.replace(model, **x) normalises the quaternion before anything is computed
        # from it (api/data.py:441-447); the kernel does the same on entry
        return ode.system_dynamics(model, x, link_forces_inertial=link_forces_inertial, joint_torques=joint_torques)

    x0 = dict(
        base_position=data._base_position, base_quaternion=q, joint_positions=data._joint_positions,
        base_linear_velocity=data._base_linear_velocity, base_angular_velocity=data._base_angular_velocity,
        joint_velocities=data._joint_velocities, contact_state=dict(data.contact_state),
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
    """``js.model.step`` with ``IntegratorType.RungeKutta4`` (``api/model.py:2601-2681``)."""
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
    tau = ode.compute_resultant_torques(model, data, joint_force_references=joint_force_references)
    return rk4_integration(model, data, fext, tau)
