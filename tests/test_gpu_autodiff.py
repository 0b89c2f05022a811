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
    # VJP (b200sim_step_vjp: the columns contracted on the device) == J^T cotangent
    ct = torch.randn(B, n_out, dtype=torch.float64, device=cuda_device)
    g = autodiff.step_vjp(model, pd, ct, joint_force_references=torch.as_tensor(tau, device=cuda_device))
    ref = torch.einsum("bo,boi->bi", ct, J)
    assert torch.allclose(g["joint_positions"], ref[:, :23], rtol=1e-10, atol=1e-12)
    assert torch.allclose(g["link_masses"], ref[:, 23:], rtol=1e-10, atol=1e-12)
    only = autodiff.step_vjp(model, pd, ct, ("link_masses",), joint_force_references=torch.as_tensor(tau, device=cuda_device))
    assert set(only) == {"link_masses"} and torch.equal(only["link_masses"], g["link_masses"])
    # another input set goes through the Jacobian
    gv = autodiff.step_vjp(model, pd, ct, ("joint_velocities",), joint_force_references=torch.as_tensor(tau, device=cuda_device))
    _, Jv, _ = autodiff.step_jacobian(model, pd, ("joint_velocities",), joint_force_references=torch.as_tensor(tau, device=cuda_device))
    assert torch.allclose(gv["joint_velocities"], torch.einsum("bo,boi->bi", ct, Jv))


def test_step_vjp_gradient_of_a_scalar_loss(cuda_device):
    """The use the reference makes of reverse mode (tests/test_automatic_differentiation.py:346-420): the gradient of a
    scalar function of the stepped state w.r.t. joint positions and link masses, here at 512 environments and for models
    with and without collidable points, against central finite differences of the C oracle."""
    import torch

    from jaxsim_b200.api import autodiff

    for name, in_contact in (("icub_like", True), ("double_pendulum", False)):
        model = H.build_model(name)
        om = H.oracle_model(model)
        n, nL = om.dofs(), om.number_of_links()
        B = 512
        od = O.random_model_data(om, B, seed=5, in_contact=in_contact)
        rng = np.random.default_rng(4)
        tau = rng.uniform(-1, 1, size=(B, n))
        w = rng.uniform(-1, 1, size=_flat(CO.step(om, od, joint_force_references=tau, caches=False)).shape)

        def loss(model_, data_):
            return float((w * _flat(CO.step(model_, data_, joint_force_references=tau, caches=False))).sum())

        pd = H.to_product(model, od, torch.float64, cuda_device)
        g = autodiff.step_vjp(model, pd, torch.as_tensor(w, device=cuda_device), joint_force_references=torch.as_tensor(tau, device=cuda_device))
        assert g["joint_positions"].shape == (B, n) and g["link_masses"].shape == (B, nL)
        eps = 1e-6
        gm = g["link_masses"].sum(dim=0).cpu().numpy()  # the masses are shared by the batch
        for k in range(0, nL, max(1, nL // 4)):
            omp, omm = copy.deepcopy(om), copy.deepcopy(om)
            omp.kin_dyn_parameters.link_parameters.mass[k] += eps
            omm.kin_dyn_parameters.link_parameters.mass[k] -= eps
            fd = (loss(omp, od) - loss(omm, od)) / (2 * eps)
            assert abs(gm[k] - fd) <= 5e-5 * max(abs(fd), np.abs(gm).max(), 1e-3), (name, "mass", k, gm[k], fd)
        gs = g["joint_positions"].cpu().numpy()
        for j in range(0, n, max(1, n // 4)):
            dp, dmn = copy.deepcopy(od), copy.deepcopy(od)
            dp.joint_positions[:, j] += eps
            dmn.joint_positions[:, j] -= eps
            # every environment is perturbed at once: the loss separates over environments
            fp = (w * _flat(CO.step(om, dp, joint_force_references=tau, caches=False))).sum(axis=1)
            fm = (w * _flat(CO.step(om, dmn, joint_force_references=tau, caches=False))).sum(axis=1)
            fd = (fp - fm) / (2 * eps)
            assert np.abs(gs[:, j] - fd).max() <= 5e-5 * max(np.abs(fd).max(), 1e-3), (name, "q", j)


def test_jvp_full_batch_properties(cuda_device):
    """BASELINE config 5 at its full size (4096 environments, float64): size-independent properties of the derivative --
    linearity in the tangent, the primal equals `step`, and the batched Jacobian columns (one launch over replicas of the
    batch) equal single-direction JVPs; finite differences of the C oracle on a strided sample."""
    import torch

    from jaxsim_b200.api import autodiff

    model = H.build_model("icub_like")
    om = H.oracle_model(model)
    B, n, nL = 4096, 23, 24
    od = O.random_model_data(om, B, seed=123, in_contact=True)
    rng = np.random.default_rng(9)
    tau = 2 * rng.uniform(-1, 1, size=(B, n))
    pd = H.to_product(model, od, torch.float64, cuda_device)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=cuda_device)  # noqa: E731
    t1 = {"joint_positions": t(rng.uniform(-1, 1, (B, n))), "link_masses": t(rng.uniform(0, 1, nL))}
    t2 = {"joint_positions": t(rng.uniform(-1, 1, (B, n))), "link_masses": t(rng.uniform(0, 1, nL))}
    a, b = 0.7, -1.3
    t3 = {k: a * t1[k] + b * t2[k] for k in t1}
    kw = dict(joint_force_references=t(tau), update_caches=False)
    p1, d1 = js.model.step_jvp(model, pd, t1, **kw)
    _, d2 = js.model.step_jvp(model, pd, t2, **kw)
    _, d3 = js.model.step_jvp(model, pd, t3, **kw)
    assert d1._link_transforms is None and p1._link_transforms is None
    ref = js.model.step(model, pd, joint_force_references=t(tau))
    for _, leaf in H.LEAVES[:6]:
        assert torch.allclose(getattr(p1, leaf), getattr(ref, leaf), rtol=1e-12, atol=1e-14), leaf
        lin = a * getattr(d1, leaf) + b * getattr(d2, leaf)
        scale = float(lin.abs().max()) + 1e-30
        assert float((getattr(d3, leaf) - lin).abs().max()) <= 1e-10 * scale, leaf
    # batched Jacobian columns == single directions; finite differences on a strided sample
    sample = torch.arange(0, B, 512, device=cuda_device)
    from jaxsim_b200.api.data import _map_leaves
    ps = _map_leaves(pd, lambda x: x[sample].contiguous())
    _, J, layout = autodiff.step_jacobian(model, ps, ("joint_positions", "link_masses"), joint_force_references=t(tau)[sample])
    e = torch.zeros(len(sample), n, dtype=torch.float64, device=cuda_device)
    e[:, 5] = 1.0
    _, dj = js.model.step_jvp(model, ps, {"joint_positions": e}, joint_force_references=t(tau)[sample], update_caches=False)
    assert torch.allclose(J[:, :n, 5], dj._joint_positions, rtol=1e-12, atol=1e-14)
    idx = sample.cpu().numpy()
    ods = O.data_replace(om, od.joint_positions[idx], od.joint_velocities[idx], od.base_quaternion[idx], od.base_linear_velocity[idx],
                         od.base_angular_velocity[idx], od.base_position[idx], od.tangential_deformation[idx])
    eps = 1e-6
    Jn = J.cpu().numpy()
    for j in (0, 11, 22):
        dp, dm = copy.deepcopy(ods), copy.deepcopy(ods)
        dp.joint_positions[:, j] += eps
        dm.joint_positions[:, j] -= eps
        fd = (_flat(CO.step(om, dp, joint_force_references=tau[idx], caches=False)) - _flat(CO.step(om, dm, joint_force_references=tau[idx], caches=False))) / (2 * eps)
        assert np.abs(Jn[:, :, j] - fd).max() / max(np.abs(fd).max(), 1e-3) <= 5e-5, ("q", j)
    for k in (0, 7, 23):
        omp, omm = copy.deepcopy(om), copy.deepcopy(om)
        omp.kin_dyn_parameters.link_parameters.mass[k] += eps
        omm.kin_dyn_parameters.link_parameters.mass[k] -= eps
        fd = (_flat(CO.step(omp, ods, joint_force_references=tau[idx], caches=False)) - _flat(CO.step(omm, ods, joint_force_references=tau[idx], caches=False))) / (2 * eps)
        assert np.abs(Jn[:, :, n + k] - fd).max() / max(np.abs(fd).max(), 1e-3) <= 5e-5, ("mass", k)
