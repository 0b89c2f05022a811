"""Weld constraints between pairs of frames: ``jaxsim.rbda.kinematic_constraints``
(``src/jaxsim/rbda/kinematic_constraints.py:19-345``), batched on the device.

The reference solves, per environment, ``(J M^-1 J^T + eps I) w = -(J nu_dot_free + baumgarte)`` for the 6D wrench
``w`` of every constraint and applies ``+w`` / ``-w`` on the parent links of the two frames before the forward dynamics
(``api/ode.py:75-107``).  Here the two rigid-body-dynamics ingredients are the CUDA kernels behind
``forward_dynamics_aba`` and ``free_floating_mass_matrix`` (one launch each for the whole batch); the Jacobians of the
few constrained frames, the 6 n_c x 6 n_c solve and the representation changes are batched torch ops on the same stream.
The result joins the step kernel's external link forces, so the step itself stays ONE launch of the fused kernel.

Kept from the reference, on purpose: the Jacobian is assembled body-fixed (``:80-85``) yet multiplied with the MIXED
generalised velocity / free acceleration / inverse mass matrix of the enclosing block (``:250-300``); ``J_dot nu`` is
omitted (``:296``); the scatter pairs the constraint-major wrench list with the side-major parent list
(``api/ode.py:97-107``; identical for one constraint).
"""

from __future__ import annotations

import copy

import numpy as np
import torch

from jaxsim_b200.api.common import VelRepr, adjoint_from_transform, other_representation_to_inertial


def _so3_log(R: torch.Tensor) -> torch.Tensor:
    """``Rotation.log_vee`` (``math/rotation.py:87-98``): rotation matrix -> rotation vector through the unit
    quaternion (largest-diagonal branch), small-angle series at the identity; batched."""
    m = R
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    m01, m02, m10, m12, m20, m21 = m[..., 0, 1], m[..., 0, 2], m[..., 1, 0], m[..., 1, 2], m[..., 2, 0], m[..., 2, 1]
    t0 = 1 + m00 - m11 - m22
    q0 = torch.stack([m21 - m12, t0, m10 + m01, m02 + m20], -1)
    t1 = 1 - m00 + m11 - m22
    q1 = torch.stack([m02 - m20, m10 + m01, t1, m21 + m12], -1)
    t2 = 1 - m00 - m11 + m22
    q2 = torch.stack([m10 - m01, m02 + m20, m21 + m12, t2], -1)
    t3 = 1 + m00 + m11 + m22
    q3 = torch.stack([t3, m21 - m12, m02 - m20, m10 - m01], -1)
    c_a, c_b, c_c = m22 < 0, m00 > m11, m00 < -m11
    t = torch.where(c_a, torch.where(c_b, t0, t1), torch.where(c_c, t2, t3))
    q = torch.where(c_a[..., None], torch.where(c_b[..., None], q0, q1), torch.where(c_c[..., None], q2, q3))
    q = q * (0.5 / torch.sqrt(t))[..., None]
    w, v = q[..., 0], q[..., 1:]
    n_sq = (v * v).sum(-1)
    small = n_sq < 1e-16
    ws = torch.where(w == 0, torch.ones_like(w), w)
    n = torch.sqrt(torch.where(small, torch.ones_like(n_sq), n_sq))
    f_small = 2.0 / ws - (2.0 / 3.0) * n_sq / ws**3
    f_big = 2.0 * torch.atan2(torch.where(w < 0, -n, n), w.abs()) / n
    return torch.where(small, f_small, f_big)[..., None] * v


def float64_data(model, data):
    """``data`` itself if it is float64, else a batched float64 copy of its state with the kinematics caches recomputed in
    float64 (so that frame positions, velocities and contact forces are consistent to 1e-16, not 1e-7)."""
    from jaxsim_b200.api import data as _data

    if data._base_quaternion.dtype == torch.float64:
        return data
    up = lambda t, nd: (t if t.dim() > nd else t[None]).to(torch.float64).contiguous()  # noqa: E731
    cs = {k: up(v, 2) for k, v in (data.contact_state or {}).items()}
    return _data._with_caches(model, data.velocity_representation, up(data._joint_positions, 1), up(data._joint_velocities, 1),
                              up(data._base_quaternion, 1), up(data._base_linear_velocity, 1), up(data._base_angular_velocity, 1),
                              up(data._base_position, 1), cs, normalise_q=False)


def _matvec(A: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    return torch.einsum("...ij,...j->...i", A, x)


def constraint_transforms(model, data) -> torch.Tensor:
    """``_compute_constraint_transforms_batched`` (``:19-55``): ``(B, n_c, 2, 4, 4)`` = ``[W_H_F1, W_H_F2]``."""
    kd = model.kin_dyn_parameters
    c, fp, nL = kd.constraints, kd.frame_parameters, kd.number_of_links()
    W_H_L = data.link_transforms
    L_H_F = torch.as_tensor(np.asarray(fp.transform), dtype=W_H_L.dtype, device=W_H_L.device)
    l1, l2 = list(c.parent_link_idxs_1), list(c.parent_link_idxs_2)
    f1, f2 = [i - nL for i in c.frame_idxs_1], [i - nL for i in c.frame_idxs_2]
    return torch.stack([W_H_L[..., l1, :, :] @ L_H_F[f1], W_H_L[..., l2, :, :] @ L_H_F[f2]], dim=-3)


def _link_jacobians_body(model, data, links: list[int]) -> torch.Tensor:
    """``generalized_free_floating_jacobian`` with body-fixed input and output (``api/model.py:925-1043``) of the given
    links: ``(B, len(links), 6, 6 + n)``.  Column of joint i: ``Ad(B_H_i) S_i`` (``rbda/jacobian.py:128-212``), masked by
    the support of the link, seen from the link frame."""
    kd = model.kin_dyn_parameters
    W_H_L = data.link_transforms
    dtype, dev = W_H_L.dtype, W_H_L.device
    nL, n = kd.number_of_links(), kd.number_of_joints()
    # poses relative to the base LINK (rbda/jacobian.py:166-170 starts the chain with the identity at link 0)
    B_H_L = torch.linalg.inv(W_H_L[..., 0:1, :, :]) @ W_H_L
    B_X_L = adjoint_from_transform(B_H_L)  # (B, nL, 6, 6)
    S = torch.as_tensor(np.asarray(kd.motion_subspaces), dtype=dtype, device=dev)  # (nL, 6)
    cols = _matvec(B_X_L[..., 1:, :, :], S[1:])  # (B, n, 6)
    eye = torch.eye(6, dtype=dtype, device=dev).expand(cols.shape[:-2] + (6, 6))
    B_J_full = torch.cat([eye, cols.transpose(-1, -2)], dim=-1)  # (B, 6, 6+n)
    kb = kd.support_body_array_bool
    mask = np.concatenate([np.ones((nL, 5)), kb.astype(float)], axis=1)[links]  # api/model.py:1002-1011
    mask = torch.as_tensor(mask, dtype=dtype, device=dev)
    L_X_B = adjoint_from_transform(B_H_L[..., links, :, :], inverse=True)
    return L_X_B @ (mask[:, None, :] * B_J_full[..., None, :, :])


def constraint_jacobians(model, data, W_H_pairs: torch.Tensor) -> torch.Tensor:
    """``_compute_constraint_jacobians_batched`` (``:58-122``): ``(B, n_c, 6, 6 + n)`` = ``J_F1 - J_F2``, each the
    link's body-fixed Jacobian moved to the frame's mixed representation by ``Ad(FW_H_L)``."""
    c = model.kin_dyn_parameters.constraints
    W_H_L = data.link_transforms

    def side(W_H_F, links):
        L_J = _link_jacobians_body(model, data, links)
        F_H_L = torch.linalg.inv(W_H_F) @ W_H_L[..., links, :, :]
        FW_H_F = W_H_F.clone()
        FW_H_F[..., 0:3, 3] = 0
        return adjoint_from_transform(FW_H_F @ F_H_L) @ L_J

    return side(W_H_pairs[..., 0, :, :], list(c.parent_link_idxs_1)) - side(W_H_pairs[..., 1, :, :], list(c.parent_link_idxs_2))


def compute_constraint_wrenches(model, data, *, joint_force_references: torch.Tensor | None = None,
                                link_forces_inertial: torch.Tensor | None = None, regularization: float = 1e-3) -> torch.Tensor:
    """``compute_constraint_wrenches`` (``:172-345``): the inertial-fixed wrench pairs ``(B, n_c, 2, 6)`` (``(n_c, 2, 6)``
    for unbatched data) that hold the constrained frames together, given the joint forces and the other link forces
    (external + contact, inertial-fixed)."""
    from jaxsim_b200.api import model as _model

    c = model.kin_dyn_parameters.constraints
    nk = 0 if c is None else len(c)
    q = data._base_quaternion
    if nk == 0:
        return torch.zeros(q.shape[:-1] + (0, 2, 6), dtype=q.dtype, device=q.device)
    dtype, dev = q.dtype, q.device
    if dtype != torch.float64:
        # ill-conditioned by construction (see api/model.py:_link_forces_with_constraints): solved in float64
        up = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.float64, device=dev)  # noqa: E731
        lf, tau = up(link_forces_inertial), up(joint_force_references)
        if q.dim() == 1:
            lf, tau = (None if lf is None else lf[None]), (None if tau is None else tau[None])
        W = compute_constraint_wrenches(model, float64_data(model, data), joint_force_references=tau, link_forces_inertial=lf,
                                        regularization=regularization).to(dtype)
        return W[0] if q.dim() == 1 else W
    if q.dim() == 1:  # unbatched data: a batch of one
        from jaxsim_b200.api.data import _map_leaves

        lf = None if link_forces_inertial is None else torch.as_tensor(link_forces_inertial, dtype=dtype, device=dev)[None]
        tau = None if joint_force_references is None else torch.as_tensor(joint_force_references, dtype=dtype, device=dev)[None]
        return compute_constraint_wrenches(model, _map_leaves(data, lambda t: t[None]), joint_force_references=tau,
                                           link_forces_inertial=lf, regularization=regularization)[0]
    d_in, d_mx = copy.copy(data), copy.copy(data)
    d_in.velocity_representation, d_mx.velocity_representation = VelRepr.Inertial, VelRepr.Mixed

    BW_nu = d_mx.generalized_velocity
    # free acceleration: the ABA kernel works inertial-fixed; its base acceleration is then read in Mixed (:250-262)
    W_vd, sdd = _model.forward_dynamics_aba(model, d_in, joint_forces=joint_force_references, link_forces=link_forces_inertial)
    vd_mx = _model._base_acceleration_to_active(model, d_mx, W_vd, data._base_linear_velocity, data._base_angular_velocity)
    BW_nud_free = torch.cat([vd_mx, sdd], dim=-1)
    M_inv = _model.free_floating_mass_matrix_inverse(model, d_mx)

    W_H = constraint_transforms(model, data)  # (B, nk, 2, 4, 4)
    J = constraint_jacobians(model, data, W_H)  # (B, nk, 6, 6+n)
    K_P = torch.as_tensor(c.K_P, dtype=dtype, device=dev)[:, None]
    K_D = torch.as_tensor(c.K_D, dtype=dtype, device=dev)[:, None]
    vel_err = _matvec(J, BW_nu[..., None, :])
    pos_err = W_H[..., 0, 0:3, 3] - W_H[..., 1, 0:3, 3]
    ori_err = _so3_log(W_H[..., 1, 0:3, 0:3].transpose(-1, -2) @ W_H[..., 0, 0:3, 0:3])
    baum = K_P * torch.cat([pos_err, ori_err], dim=-1) + K_D * vel_err  # (B, nk, 6), :125-169

    Js = J.reshape(J.shape[:-3] + (6 * nk, J.shape[-1]))
    A = Js @ M_inv @ Js.transpose(-1, -2) + regularization * torch.eye(6 * nk, dtype=dtype, device=dev)
    rhs = _matvec(Js, BW_nud_free) + baum.reshape(baum.shape[:-2] + (6 * nk,))
    w = torch.linalg.solve(A, -rhs[..., None])[..., 0].reshape(baum.shape)
    return torch.stack([
        other_representation_to_inertial(w, VelRepr.Mixed, W_H[..., 0, :, :], is_force=True),
        other_representation_to_inertial(-w, VelRepr.Mixed, W_H[..., 1, :, :], is_force=True)], dim=-2)


def constraint_link_forces(model, data, *, joint_torques: torch.Tensor, link_forces_inertial: torch.Tensor) -> torch.Tensor:
    """The wrench pairs scattered onto the parent links like ``api/ode.py:90-107`` does: ``(B, nL, 6)``."""
    c = model.kin_dyn_parameters.constraints
    W = compute_constraint_wrenches(model, data, joint_force_references=joint_torques, link_forces_inertial=link_forces_inertial)
    flat = W.reshape(W.shape[:-3] + (2 * len(c), 6))
    parents = torch.as_tensor(list(c.parent_link_idxs_1) + list(c.parent_link_idxs_2), dtype=torch.long, device=W.device)
    out = torch.zeros(W.shape[:-3] + (model.number_of_links(), 6), dtype=W.dtype, device=W.device)
    return out.index_add_(-2, parents, flat)
