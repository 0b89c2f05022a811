"""Derivatives of ``step`` w.r.t. state and hardware parameters (BASELINE config 5).

The reference obtains them from JAX AD (``tests/test_automatic_differentiation.py:346-420``:
``jax.jacfwd`` / ``check_grads`` of one ``step``); here they come from the forward-mode
instantiation of the step kernel (``b200sim_step_jvp``; one tangent direction per dual environment, many
directions per launch by replicating the batch):

* :func:`jaxsim_b200.api.model.step_jvp` -- one Jacobian-vector product;
* :func:`step_jacobian` -- the full Jacobian d(step outputs)/d(joint positions, link masses, ...)
  assembled from batched launches (per-environment inputs) and one launch per shared model parameter;
* :func:`step_vjp` -- a vector-Jacobian product (the "gradient" of a scalar loss), assembled
  from the same columns.  Cost = number of input coordinates x one JVP launch, which is the
  right trade for the low-dimensional parameter sets of hardware co-design (n + nL inputs).
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
    batch (K x B dual environments), which keeps the forward-mode kernel in its throughput regime; a mass direction
    changes the (shared) model constants and takes a launch of its own.  Without ``with_caches`` the kernel skips the
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
        if name == "link_masses":
            for j in range(k):
                e = torch.zeros(k, dtype=torch.float64)
                e[j] = 1.0
                po, dout = _model.step_jvp(model, data, {name: e}, joint_force_references=tau, update_caches=with_caches)
                out = po if out is None else out
                cols.append(_flatten_outputs(dout, with_caches))
            continue
        Kmax = max(1, min(k, _MAX_DUAL_ENVS // max(B, 1)))
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


def step_vjp(model, data, cotangent: torch.Tensor, wrt=("joint_positions", "link_masses"), *,
             joint_force_references=None) -> dict:
    """Vector-Jacobian product: ``cotangent`` is ``(B, n_out)`` over the flattened new state
    (same ordering as :func:`step_jacobian`, no caches).  Returns ``{name: gradient}`` with
    per-environment gradients ``(B, size)`` (for ``link_masses`` too: sum over the batch for
    the gradient of a batch-summed loss w.r.t. the shared masses)."""
    _, J, layout = step_jacobian(model, data, wrt, joint_force_references=joint_force_references)
    g = torch.einsum("bo,boi->bi", cotangent.to(J.dtype), J)
    return {name: g[:, sl] for name, sl in layout.items()}
