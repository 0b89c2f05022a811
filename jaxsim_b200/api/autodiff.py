"""Derivatives of ``step`` w.r.t. state and hardware parameters (BASELINE config 5).

The reference obtains them from JAX AD (``tests/test_automatic_differentiation.py:346-420``:
``jax.jacfwd`` / ``check_grads`` of one ``step``); here they come from the forward-mode
instantiation of the step kernel (``b200sim_step_jvp``; one tangent direction per dual environment, many
directions per launch by replicating the batch):

* :func:`jaxsim_b200.api.model.step_jvp` -- one Jacobian-vector product;
* :func:`step_jacobian` -- the full Jacobian d(step outputs)/d(joint positions, link masses, ...)
  assembled from batched launches (per-environment inputs) and one launch per shared model parameter;
* :func:`step_vjp` -- a vector-Jacobian product (the "gradient" of a scalar loss): one call of
  ``b200sim_step_vjp``, which contracts the same forward-mode columns on the device.  Cost = about one forward-mode
  step per input coordinate, the right trade for the low-dimensional parameter sets of hardware co-design (n + nL).
"""

from __future__ import annotations

import copy

import torch

from . import model as _model

STATE_LEAVES = (
    "joint_positions", "joint_velocities", "base_quaternion", "base_linear_velocity",
    "base_angular_velocity", "base_position",
)


def _leaf_sizes(model) -> dict:
    n = model.dofs()
    return {"joint_positions": n, "joint_velocities": n, "base_quaternion": 4, "base_linear_velocity": 3,
            "base_angular_velocity": 3, "base_position": 3, "link_masses": model.number_of_links(),
            "joint_force_references": n}


def _flatten_outputs(d, with_caches: bool) -> torch.Tensor:
    B = d._base_quaternion.shape[0]
    parts = [d._joint_positions, d._joint_velocities, d._base_quaternion, d._base_linear_velocity,
             d._base_angular_velocity, d._base_position]
    if "tangential_deformation" in d.contact_state:
        parts.append(d.contact_state["tangential_deformation"].reshape(B, -1))
    if with_caches:
        parts += [d._base_transform.reshape(B, -1), d._joint_transforms.reshape(B, -1),
                  d._link_transforms.reshape(B, -1), d._link_velocities.reshape(B, -1)]
    return torch.cat([p.reshape(B, -1) for p in parts], dim=-1)


# environments x directions per launch: enough to fill the GPU several times over, small enough for the dual caches
_MAX_DUAL_ENVS = 131072


def step_jacobian(model, data, wrt=("joint_positions", "link_masses"), *, joint_force_references=None,
                  with_caches: bool = False):
    """Jacobian of one ``step`` w.r.t. the inputs named in ``wrt``.

    Returns ``(data_out, J, layout)``: ``J`` has shape ``(B, n_out, n_in)`` where the output
    coordinates are the flattened new state leaves (joint positions, joint velocities, base
    quaternion, base linear / angular velocity, base position, tangential deformation
    [, caches]) and the input coordinates are the concatenation of the ``wrt`` leaves;
    ``layout`` maps each ``wrt`` name to its column slice.  ``link_masses`` is the shared
    ``LinkParameters.mass`` vector (``api/kin_dyn_parameters.py:596``).

    Directions of per-environment inputs are batched: K unit directions run as ONE launch over K replicas of the
    batch (K x B dual environments), which keeps the forward-mode kernel in its throughput regime -- link-mass directions
    too (``b200sim_step_jvp_ex``: the constants carry the unit direction of every link, each replica keeps one).  Without ``with_caches`` the kernel skips the
    kinematics of the new state altogether."""
    from .data import _map_leaves

    sizes = _leaf_sizes(model)
    q = data._base_quaternion
    B, dev = q.shape[0], q.device
    tau = None if joint_force_references is None else torch.as_tensor(joint_force_references, dtype=torch.float64, device=dev).expand(B, model.dofs())
    cols, layout, o = [], {}, 0
    out = None
    replicas = {}  # K -> (data replicated K times, tau replicated)

    def replicated(K):
        if K not in replicas:
            rep = lambda t: t.repeat((K,) + (1,) * (t.dim() - 1))  # noqa: E731
            bare = copy.copy(data)  # the forward-mode kernel reads the state leaves only: do not replicate the caches
            bare._base_transform = bare._joint_transforms = bare._link_transforms = bare._link_velocities = None
            replicas[K] = (_map_leaves(bare, rep), None if tau is None else rep(tau))
        return replicas[K]

    for name in wrt:
        k = sizes[name]
        layout[name] = slice(o, o + k)
        o += k
        Kmax = max(1, min(k, _MAX_DUAL_ENVS // max(B, 1)))
        if name == "link_masses":
            # the constants carry the unit direction of every link; replica r keeps it for link j0 + r alone
            for j0 in range(0, k, Kmax):
                K = min(Kmax, k - j0)
                dK, tauK = replicated(K)
                po, dout = _model.step_jvp(model, dK, {name: torch.ones(k, dtype=torch.float64)}, joint_force_references=tauK,
                                           update_caches=with_caches, mass_direction_period=B, mass_direction_first_link=j0)
                if out is None:
                    out = _map_leaves(po, lambda t: t[:B])
                flat = _flatten_outputs(dout, with_caches).reshape(K, B, -1)
                cols.extend(flat[i] for i in range(K))
            continue
        for j0 in range(0, k, Kmax):
            K = min(Kmax, k - j0)
            dK, tauK = replicated(K)
            e = torch.zeros(K, B, k, dtype=torch.float64, device=dev)
            e[torch.arange(K, device=dev), :, j0 + torch.arange(K, device=dev)] = 1.0
            po, dout = _model.step_jvp(model, dK, {name: e.reshape(K * B, k)}, joint_force_references=tauK, update_caches=with_caches)
            if out is None:
                out = _map_leaves(po, lambda t: t[:B])
            flat = _flatten_outputs(dout, with_caches).reshape(K, B, -1)
            cols.extend(flat[i] for i in range(K))
    J = torch.stack(cols, dim=-1) if cols else torch.zeros(B, 0, 0, dtype=torch.float64, device=dev)
    return out, J, layout


def step_vjp(model, data, cotangent, wrt=("joint_positions", "link_masses"), *,
             joint_force_references=None) -> dict:
    """Vector-Jacobian product of one ``step``: ``cotangent`` is ``(B, n_out)`` over the flattened new state (same
    ordering as :func:`step_jacobian`, no caches) or a ``JaxSimModelData``-like object / dict of cotangent leaves.
    Returns ``{name: gradient}`` with per-environment gradients ``(B, size)`` (for ``link_masses`` too: sum over the
    batch for the gradient of a batch-summed loss w.r.t. the shared masses).

    For ``wrt`` within (joint positions, link masses) -- BASELINE config 5 -- this is ONE call of ``b200sim_step_vjp``:
    forward-mode columns contracted on the device.  Other inputs go through :func:`step_jacobian`."""
    from jaxsim_b200 import _lib
    from jaxsim_b200.rbda.contacts import SoftContacts

    q = data._base_quaternion
    if q.dim() != 2 or q.dtype != torch.float64:
        raise ValueError("step_vjp needs batched float64 data")
    B, dev = q.shape[0], q.device
    n, nL, nc = model.dofs(), model.number_of_links(), model.number_of_collidable_points()
    soft = isinstance(model.contact_model, SoftContacts) and nc > 0
    widths = [("s", n), ("sd", n), ("q", 4), ("vl", 3), ("om", 3), ("p", 3)] + ([("m", 3 * nc)] if soft else [])
    ct = torch.as_tensor(cotangent, dtype=torch.float64, device=dev)
    if ct.shape != (B, sum(w for _, w in widths)):
        raise ValueError(ct.shape, (B, sum(w for _, w in widths)))
    if not set(wrt) <= {"joint_positions", "link_masses"}:
        _, J, layout = step_jacobian(model, data, wrt, joint_force_references=joint_force_references)
        g = torch.einsum("bo,boi->bi", ct, J)
        return {name: g[:, sl] for name, sl in layout.items()}
    parts, o = {}, 0
    for name, w in widths:
        parts[name] = ct[:, o:o + w].contiguous()
        o += w
    c = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.float64, device=dev).contiguous()  # noqa: E731
    m = c(data.contact_state.get("tangential_deformation")) if (soft and data.contact_state) else None
    tau = None if joint_force_references is None else c(torch.as_tensor(joint_force_references, dtype=torch.float64, device=dev).expand(B, n))
    g_s = torch.empty(B, n, dtype=torch.float64, device=dev) if "joint_positions" in wrt else None
    g_m = torch.empty(B, nL, dtype=torch.float64, device=dev) if "link_masses" in wrt else None
    dm = model.device_model(dev)
    ptr, stream = _model._ptr, _model._stream_ptr(dev)
    with torch.cuda.device(dev):
        rc = _lib.load().b200sim_step_vjp(
            dm.handle, B, ptr(c(data._joint_positions)), ptr(c(data._joint_velocities)), ptr(c(q)), ptr(c(data._base_linear_velocity)),
            ptr(c(data._base_angular_velocity)), ptr(c(data._base_position)), ptr(m), ptr(tau),
            ptr(parts["s"]), ptr(parts["sd"]), ptr(parts["q"]), ptr(parts["vl"]), ptr(parts["om"]), ptr(parts["p"]), ptr(parts.get("m")),
            ptr(g_s), ptr(g_m), stream)
    _lib.check(rc, "b200sim_step_vjp")
    out = {}
    for name in wrt:
        out[name] = g_s if name == "joint_positions" else g_m
    return out
