"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same inputs."""

import numpy as np
import pytest

import jaxsim_b200.api as js
from oracle import jaxsim_oracle as O

from . import helpers as H

pytestmark = pytest.mark.gpu

MODELS = ["pendulum", "double_pendulum", "cartpole", "box", "sphere", "icub_like", "ergocub_like"]


def _dtype(name):
    import torch

    return {"float64": torch.float64, "float32": torch.float32}[name]


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", MODELS)
def test_fk_caches(name, dtype, cuda_device):
    """JaxSimModelData.build caches == oracle data_replace (api/data.py:66-202)."""
    model = H.build_model(name)
    om = H.oracle_model(model)
    od = O.random_model_data(om, 33, seed=3)
    pd = H.to_product(model, od, _dtype(dtype), cuda_device)
    H.compare_data(pd, od, H.RTOL[dtype], f"fk {name} {dtype}")


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", MODELS)
def test_aba(name, dtype, cuda_device):
    """forward_dynamics_aba == oracle aba (rbda/aba.py) with random torques + link forces."""
    import torch

    model = H.build_model(name)
    om = H.oracle_model(model)
    B = 37
    od = O.random_model_data(om, B, seed=5)
    rng = np.random.default_rng(0)
    tau = 10 * rng.uniform(size=(B, om.dofs()))
    W_f = rng.uniform(size=(B, om.number_of_links(), 6))
    a_ref, sdd_ref = O.aba(om, od.base_position, od.base_orientation, od.joint_positions, od.base_linear_velocity,
                           od.base_angular_velocity, od.joint_velocities, tau, W_f)
    pd = H.to_product(model, od, _dtype(dtype), cuda_device)
    t = lambda a: torch.as_tensor(a, dtype=_dtype(dtype), device=cuda_device)  # noqa: E731
    a, sdd = js.model.forward_dynamics_aba(model, pd, joint_forces=t(tau), link_forces=t(W_f))
    rt = H.RTOL[dtype]
    assert H.rel_err(a.cpu().numpy(), a_ref) <= rt
    if om.dofs():
        assert H.rel_err(sdd.cpu().numpy(), sdd_ref) <= rt


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("in_contact", [False, True])
@pytest.mark.parametrize("name", MODELS)
def test_step(name, in_contact, dtype, cuda_device):
    """js.model.step == oracle step (api/model.py:2601-2681), every output leaf."""
    import torch

    model = H.build_model(name)
    om = H.oracle_model(model)
    B = 41
    od = O.random_model_data(om, B, seed=7, in_contact=in_contact)
    rng = np.random.default_rng(1)
    tau = 10 * rng.uniform(size=(B, om.dofs()))
    od.tangential_deformation = 1e-4 * rng.uniform(-1, 1, size=od.tangential_deformation.shape)
    od.tangential_deformation[..., 2] = 0.0
    ref = O.step(om, od, joint_force_references=tau)
    pd = H.to_product(model, od, _dtype(dtype), cuda_device)
    out = js.model.step(model, pd, joint_force_references=torch.as_tensor(tau, dtype=_dtype(dtype), device=cuda_device))
    H.compare_data(out, ref, H.RTOL[dtype], f"step {name} contact={in_contact} {dtype}")


@pytest.mark.parametrize("G", [1, 2, 4, 8, 16, 32])
def test_step_all_lane_widths(G, cuda_device):
    """Results do not depend on the lanes-per-environment tuning knob."""
    import torch

    model = H.build_model("icub_like")
    model.set_tuning(lanes_per_env=G)
    om = H.oracle_model(model)
    B = 50
    od = O.random_model_data(om, B, seed=11, in_contact=True)
    ref = O.step(om, od)
    pd = H.to_product(model, od, torch.float64, cuda_device)
    out = js.model.step(model, pd)
    H.compare_data(out, ref, 1e-5, f"G={G}")


def test_step_link_forces_representations(cuda_device):
    """link_forces are interpreted in data.velocity_representation (api/model.py:2641-2646)."""
    import torch

    model = H.build_model("icub_like")
    om = H.oracle_model(model)
    B = 9
    od = O.random_model_data(om, B, seed=13)
    rng = np.random.default_rng(2)
    O_f = rng.uniform(-5, 5, size=(B, om.number_of_links(), 6))
    for vr, name in ((js.common.VelRepr.Inertial, "inertial"), (js.common.VelRepr.Body, "body"), (js.common.VelRepr.Mixed, "mixed")):
        W_f = O.other_representation_to_inertial(O_f, name, od.link_transforms, is_force=True)
        ref = O.step(om, od, link_forces_inertial=W_f)
        pd = H.to_product(model, od, torch.float64, cuda_device, velocity_representation=vr)
        out = js.model.step(model, pd, link_forces=torch.as_tensor(O_f, device=cuda_device))
        assert out.velocity_representation == vr
        H.compare_data(out, ref, 1e-5, f"link_forces {name}")


def test_rollout_matches_oracle(cuda_device):
    """100 consecutive steps of a falling, landing box stay within tolerance (fp64)."""
    import torch

    model = H.build_model("box")
    om = H.oracle_model(model)
    B = 4
    od = O.random_model_data(om, B, seed=17, base_pos_bounds=((-1, -1, 0.06), (1, 1, 0.12)))
    pd = H.to_product(model, od, torch.float64, cuda_device)
    for _ in range(100):
        od = O.step(om, od)
        pd = js.model.step(model, pd)
    H.compare_data(pd, od, 1e-5, "box rollout")
