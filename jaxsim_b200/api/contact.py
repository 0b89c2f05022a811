"""Host-side helpers of ``jaxsim.api.contact`` that sit next to the hot path: collidable-point
kinematics read from the cached link transforms / velocities (``api/contact.py:18-45``,
``rbda/collidable_points.py:9-65``) and ``estimate_good_contact_parameters`` (``:141-211``).
Small batched torch ops on the device; the contact MODELS live in the CUDA kernels."""

from __future__ import annotations

import numpy as np
import torch

STANDARD_GRAVITY = 9.81  # src/jaxsim/math/__init__.py:14

MAX_STIFFNESS = 1e6  # rbda/contacts/common.py:21-22
MAX_DAMPING = 1e4


def _enabled(model, device):
    cp = model.kin_dyn_parameters.contact_parameters
    idx = [k for k, e in enumerate(cp.enabled) if e]
    body = torch.as_tensor(np.asarray(cp.body, dtype=np.int64)[idx], device=device)
    point = np.asarray(cp.point, dtype=np.float64).reshape(-1, 3)[idx]
    return body, point


def collidable_point_kinematics(model, data) -> tuple[torch.Tensor, torch.Tensor]:
    """``js.contact.collidable_point_kinematics``: position and (mixed) linear velocity of the
    enabled collidable points, ``(B, nc, 3)`` each, from the data's cached link transforms and
    velocities: ``W_p_C = W_H_L [L_p_C; 1]``, ``p_dot_C = v_lin + omega x W_p_C``."""
    H = data.link_transforms
    V = data.link_velocities
    body, point = _enabled(model, H.device)
    Lp = torch.as_tensor(point, dtype=H.dtype, device=H.device)
    Hb = H[..., body, :, :]
    W_p = torch.einsum("...kij,kj->...ki", Hb[..., 0:3, 0:3], Lp) + Hb[..., 0:3, 3]
    Vb = V[..., body, :]
    W_pd = Vb[..., 0:3] + torch.linalg.cross(Vb[..., 3:6], W_p)
    return W_p, W_pd


def collidable_point_positions(model, data) -> torch.Tensor:
    return collidable_point_kinematics(model, data)[0]


def collidable_point_velocities(model, data) -> torch.Tensor:
    return collidable_point_kinematics(model, data)[1]


def in_contact(model, data) -> torch.Tensor:
    """``js.contact.in_contact`` (``api/contact.py:83-143``) for all links: ``(B, nL)`` booleans, a
    link is in contact when one of its enabled points lies at or below the (flat) terrain."""
    W_p = collidable_point_positions(model, data)
    body, _ = _enabled(model, W_p.device)
    below = W_p[..., 2] <= model.terrain.height()
    onehot = torch.nn.functional.one_hot(body, model.number_of_links()).to(torch.bool)  # (nc, nL)
    return (below[..., :, None] & onehot).any(dim=-2)


def com_position(model, data) -> torch.Tensor:
    """``js.com.com_position`` (``api/com.py``): ``sum_i m_i W_H_i L_p_com_i / m``."""
    lp = model.kin_dyn_parameters.link_parameters
    H = data.link_transforms
    m = torch.as_tensor(np.asarray(lp.mass), dtype=H.dtype, device=H.device)
    c = torch.as_tensor(np.asarray(lp.center_of_mass), dtype=H.dtype, device=H.device)
    W_c = torch.einsum("...kij,kj->...ki", H[..., 0:3, 0:3], c) + H[..., 0:3, 3]
    return (m[:, None] * W_c).sum(dim=-2) / m.sum()


def estimate_good_contact_parameters(model, *, standard_gravity: float = STANDARD_GRAVITY,
                                     static_friction_coefficient: float = 0.5,
                                     number_of_active_collidable_points_steady_state: int = 1,
                                     damping_ratio: float = 1.0, max_penetration: float | None = None,
                                     device: torch.device | str = "cuda"):
    """``js.contact.estimate_good_contact_parameters`` (``api/contact.py:155-211`` +
    ``ContactsParams.build_default_from_jaxsim_model``, ``rbda/contacts/common.py:88-168``): the
    stiffness that gives ``max_penetration`` at steady state on ``nc`` points (Hunt/Crossley
    exponent p = 0.5), the damping from the damping ratio, both clipped; returned as the
    parameter class of the model's contact model."""
    from . import data as _data

    if max_penetration is None:
        zero = _data.JaxSimModelData.build(model, device=torch.device(device), dtype=torch.float64)
        z_com = float(com_position(model, zero)[..., 2].reshape(-1)[0])
        if model.floating_base() and model.number_of_collidable_points() > 0:
            z_com -= float(collidable_point_positions(model, zero)[..., 2].min())
        max_penetration = 0.01 * z_com  # 1 % of the centre-of-mass height
    m = float(np.asarray(model.kin_dyn_parameters.link_parameters.mass).sum())
    p = q = 0.5
    f_average = m * standard_gravity / number_of_active_collidable_points_steady_state
    stiffness = float(np.clip(f_average / max_penetration ** (1 + p), 0, MAX_STIFFNESS))
    damping = float(np.clip(damping_ratio * 2 * np.sqrt(stiffness * m), 0, MAX_DAMPING))
    return model.contact_model._parameters_class.build(K=stiffness, D=damping, mu=static_friction_coefficient, p=p, q=q)


def soft_link_contact_forces(model, data) -> tuple[torch.Tensor, torch.Tensor]:
    """``js.contact.link_contact_forces`` for ``SoftContacts`` (``api/contact.py:516-554``): the inertial-fixed 6D
    contact force on every link, ``(B, nL, 6)``, and the derivative of the tangential deformation, ``(B, nc, 3)``
    (zero for disabled points), from the data's cached link transforms / velocities.  Hunt/Crossley with flat terrain
    (``rbda/contacts/soft.py:195-444``) as batched torch ops: used by the RungeKutta4Fast integrator, which evaluates the
    contact model once per step OUTSIDE the dynamics kernel (``api/integrators.py:175-188``); every other path evaluates
    the same model inside the CUDA kernels."""
    prm = model.contact_params
    cp = model.kin_dyn_parameters.contact_parameters
    W_p, W_pd = collidable_point_kinematics(model, data)
    dtype, dev = W_p.dtype, W_p.device
    idx = torch.as_tensor([k for k, e in enumerate(cp.enabled) if e], dtype=torch.long, device=dev)
    m_all = data.contact_state["tangential_deformation"]
    m = m_all[..., idx, :]
    K, D, mu = float(prm.K), float(prm.D), float(prm.mu)
    eps = torch.finfo(dtype).eps
    delta = torch.clamp(model.terrain.height() - W_p[..., 2], min=0.0)
    ddot = torch.where(delta > 0, -W_pd[..., 2], torch.zeros_like(delta))
    dp = torch.pow(delta + eps, float(prm.p))
    dq = torch.pow(delta + eps, float(prm.q))
    fn = torch.clamp((K * dp) * delta + (D * dq) * ddot, min=0.0)
    zero = torch.zeros_like(delta)
    v_t = torch.stack([W_pd[..., 0], W_pd[..., 1], zero], dim=-1)
    m_n = torch.stack([zero, zero, m[..., 2]], dim=-1)
    m_t = torch.stack([m[..., 0], m[..., 1], zero], dim=-1)
    f_t = -((K * dp)[..., None] * m_t + (D * dq)[..., None] * v_t)
    sticking = (delta <= 0) | ((f_t * f_t).sum(-1) <= (mu * fn) ** 2)
    nrm = torch.linalg.norm(f_t, dim=-1)
    direction = f_t / (nrm + eps * (nrm == 0))[..., None]
    f_t = torch.where(sticking[..., None], f_t, torch.minimum(mu * fn, nrm)[..., None] * direction)
    f_t = torch.where((delta <= 0)[..., None], torch.zeros_like(f_t), f_t)
    md_no = -(K / D) * m
    md_stick = v_t - (K / D) * m_n
    md_slip = -(f_t + (K * dp)[..., None] * m_t) / (D * dq)[..., None]
    status = sticking.to(torch.int64) + (delta <= 0).to(torch.int64)
    md = torch.where((status == 0)[..., None], md_slip, torch.where((status == 1)[..., None], md_stick, md_no))
    f = f_t + torch.stack([zero, zero, fn], dim=-1)
    W_f_C = torch.cat([f, torch.linalg.cross(W_p, f)], dim=-1)  # W_f = [f; p x f] (soft.py:378-386)
    body, _ = _enabled(model, dev)
    W_f_L = torch.zeros(W_f_C.shape[:-2] + (model.number_of_links(), 6), dtype=dtype, device=dev)
    W_f_L.index_add_(-2, body, W_f_C)
    m_dot = torch.zeros_like(m_all)
    m_dot[..., idx, :] = md
    return W_f_L, m_dot
