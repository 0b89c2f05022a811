"""Derivatives of ``step`` w.r.t. state and hardware parameters (BASELINE config 5).

The reference obtains them from JAX AD (``tests/test_automatic_differentiation.py:346-420``:
``jax.jacfwd`` / ``check_grads`` of one ``step``); here they come from the forward-mode
instantiation of the step kernel (``b200sim_step_jvp``, one tangent direction per launch):

* :func:`jaxsim_b200.api.model.step_jvp` -- one Jacobian-vector product;
* :func:`step_jacobian` -- the full Jacobian d(step outputs)/d(joint positions, link masses, ...)
  assembled column by column (one launch per input coordinate);
* :func:`step_vjp` -- a vector-Jacobian product (the "gradient" of a scalar loss), assembled
  from the same columns.  Cost = number of input coordinates x one JVP launch, which is the
  right trade for the low-dimensional parameter sets of hardware co-design (n + nL inputs).
"""

from __future__ import annotations

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


def step_jacobian(model, data, wrt=("joint_positions", "link_masses"), *, joint_force_references=None,
                  with_caches: bool = False):
    """Jacobian of one ``step`` w.r.t. the inputs named in ``wrt``.

    Returns ``(data_out, J, layout)``: ``J`` has shape ``(B, n_out, n_in)`` where the output
    coordinates are the flattened new state leaves (joint positions, joint velocities, base
    quaternion, base linear / angular velocity, base position, tangential deformation
    [, caches]) and the input coordinates are the concatenation of the ``wrt`` leaves;
    ``layout`` maps each ``wrt`` name to its column slice.  ``link_masses`` is the shared
    ``LinkParameters.mass`` vector (``api/kin_dyn_parameters.py:596``)."""
    sizes = _leaf_sizes(model)
    q = data._base_quaternion
    B, dev = q.shape[0], q.device
    cols, layout, o = [], {}, 0
    out = None
    for name in wrt:
        k = sizes[name]
        layout[name] = slice(o, o + k)
        o += k
        for j in range(k):
            if name == "link_masses":
                e = torch.zeros(k, dtype=torch.float64)
                e[j] = 1.0
            else:
                e = torch.zeros(B, k, dtype=torch.float64, device=dev)
                e[:, j] = 1.0
            out, dout = _model.step_jvp(model, data, {name: e}, joint_force_references=joint_force_references)
            cols.append(_flatten_outputs(dout, with_caches))
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
