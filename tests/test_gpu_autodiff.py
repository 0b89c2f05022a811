"""-m gpu: BASELINE config 5 -- d(step)/d(link masses, joint positions) for the iCub-like
model in float64: the full Jacobian assembled from forward-mode launches, checked against
central finite differences of the fp64 oracle (the reference validates its AD with
jax.test_util.check_grads the same way, tests/test_automatic_differentiation.py:24-27)."""

import copy

import numpy as np
import pytest

import jaxsim_b200.api as js
from oracle import c_oracle as CO
from oracle import jaxsim_oracle as O

from . import helpers as H

pytestmark = pytest.mark.gpu


def _flat(d: O.OracleData) -> np.ndarray:
    B = d.base_quaternion.shape[0]
    return np.concatenate([d.joint_positions, d.joint_velocities, d.base_quaternion, d.base_linear_velocity,
                           d.base_angular_velocity, d.base_position, d.tangential_deformation.reshape(B, -1)], axis=1)


def test_step_jacobian_wrt_masses_and_joint_positions(cuda_device):
    import torch

    from jaxsim_b200.api import autodiff

    model = H.build_model("icub_like")
    om = H.oracle_model(model)
    B = 6
    od = O.random_model_data(om, B, seed=97, in_contact=True)
    tau = 2 * np.random.default_rng(1).uniform(-1, 1, size=(B, 23))
    pd = H.to_product(model, od, torch.float64, cuda_device)
    out, J, layout = autodiff.step_jacobian(model, pd, ("joint_positions", "link_masses"),
                                            joint_force_references=torch.as_tensor(tau, device=cuda_device))
    n_out = 23 + 23 + 4 + 3 + 3 + 3 + 16 * 3
    assert J.shape == (B, n_out, 23 + 24) and layout["link_masses"] == slice(23, 47)
    H.compare_data(out, O.step(om, od, joint_force_references=tau), 1e-9, "jacobian primal")
    Jn = J.cpu().numpy()
    eps = 1e-6
    # finite differences with the C oracle (fast): joint positions
    for j in range(23):
        dp, dm = copy.deepcopy(od), copy.deepcopy(od)
        dp.joint_positions[:, j] += eps
        dm.joint_positions[:, j] -= eps
        fd = (_flat(CO.step(om, dp, joint_force_references=tau, caches=False)) - _flat(CO.step(om, dm, joint_force_references=tau, caches=False))) / (2 * eps)
        scale = max(np.abs(fd).max(), 1e-3)
        assert np.abs(Jn[:, :, j] - fd).max() / scale <= 5e-5, ("q", j)
    # link masses
    for k in range(24):
        omp, omm = copy.deepcopy(om), copy.deepcopy(om)
        omp.kin_dyn_parameters.link_parameters.mass[k] += eps
        omm.kin_dyn_parameters.link_parameters.mass[k] -= eps
        fd = (_flat(CO.step(omp, od, joint_force_references=tau, caches=False)) - _flat(CO.step(omm, od, joint_force_references=tau, caches=False))) / (2 * eps)
        scale = max(np.abs(fd).max(), 1e-3)
        assert np.abs(Jn[:, :, 23 + k] - fd).max() / scale <= 5e-5, ("mass", k)
    # VJP == J^T cotangent
    ct = torch.randn(B, n_out, dtype=torch.float64, device=cuda_device)
    g = autodiff.step_vjp(model, pd, ct, joint_force_references=torch.as_tensor(tau, device=cuda_device))
    ref = torch.einsum("bo,boi->bi", ct, J)
    assert torch.allclose(g["joint_positions"], ref[:, :23]) and torch.allclose(g["link_masses"], ref[:, 23:])
