"""-m gpu: weld constraints (reference ``rbda/kinematic_constraints.py``, ``api/ode.py:75-107``) -- the
device-side wrench solve against the reference's own outputs and the oracle, and the reference's behavioural test
(``tests/test_simulations.py:549-612``: a welded 4-bar linkage stays closed while it is simulated)."""

import json

import numpy as np
import pytest

import jaxsim_b200.api as js
from oracle import constraints_oracle as KC
from oracle import jaxsim_oracle as O
from tests.golden import cases as C

from . import helpers as H

pytestmark = pytest.mark.gpu

WELD_CASES = [c["id"] for c in C.all_cases() if c["constraints"]]


def _fixture(cid):
    z = np.load(C.fixture_path(cid), allow_pickle=False)
    return z, json.loads(str(z["spec"]))


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("cid", WELD_CASES)
def test_constraint_wrenches_match_reference(cid, dtype, cuda_device):
    import torch

    from jaxsim_b200.rbda.kinematic_constraints import compute_constraint_wrenches

    z, _ = _fixture(cid)
    case = C.case(cid)
    model = H.build_model_for_case(case)
    om = H.oracle_model(model)
    td = torch.float64 if dtype == "float64" else torch.float32
    od = O.data_replace(om, z["in_joint_positions"], z["in_joint_velocities"], z["in_base_quaternion"],
                        z["in_base_linear_velocity"], z["in_base_angular_velocity"], z["in_base_position"],
                        z["in_tangential_deformation"])
    data = H.to_product(model, od, td, cuda_device)
    W = compute_constraint_wrenches(model, data).cpu().numpy()
    ref = z["constraint_wrenches_free"]
    assert W.shape == ref.shape
    assert H.elementwise_violation(W, ref, H.RTOL[dtype], batched=True) <= 1.0
    # frames of the constraint: js.frame.transform
    c = model.kin_dyn_parameters.constraints
    T = torch.stack([js.frame.transform(model, data, frame_index=i) for i in (c.frame_idxs_1[0], c.frame_idxs_2[0])], dim=1)
    assert H.elementwise_violation(T.cpu().numpy(), z["constraint_frame_transforms"], 1e-6 if dtype == "float32" else 1e-12, batched=True) <= 1.0
    # unbatched data gives the unbatched result
    W0 = compute_constraint_wrenches(model, _env(data, 0)).cpu().numpy()
    assert W0.shape == ref.shape[1:]
    assert H.elementwise_violation(W0[None], ref[0:1], H.RTOL[dtype], batched=True) <= 1.0


def _env(data, b):
    from jaxsim_b200.api.data import _map_leaves

    return _map_leaves(data, lambda t: t[b])


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_constrained_step_matches_oracle_large_batch(dtype, cuda_device):
    """64 random environments, torques and Body-fixed external forces: product step vs the pinned oracle.  float64 with
    the loop wide open (joint angles over their whole range), float32 with a nearly closed loop: an open loop makes the
    solve return 1e7 N wrench pairs that cancel to 1e3 N, which no float32 force array carries (tests/golden/cases.py)."""
    import torch

    case = C.case("four_bar_weld_contact")
    model = H.build_model_for_case(case)
    om = H.oracle_model(model)
    td = torch.float64 if dtype == "float64" else torch.float32
    B = 64
    od = O.random_model_data(om, B, seed=77, in_contact=True)
    rng = np.random.default_rng(5)
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)  # noqa: E731
    js_ = 1.0 if dtype == "float64" else 0.02
    od = O.data_replace(om, f32(js_ * od.joint_positions), f32(js_ * od.joint_velocities), f32(od.base_quaternion), f32(od.base_linear_velocity),
                        f32(od.base_angular_velocity), f32(od.base_position), f32(1e-4 * rng.uniform(-1, 1, (B, 8, 3))))
    tau = f32(rng.uniform(-3, 3, (B, om.dofs())))
    fb = f32(rng.uniform(-5, 5, (B, om.number_of_links(), 6)))
    W_f = O.other_representation_to_inertial(fb, "body", od.link_transforms, is_force=True)
    ref = KC.step(om, od, link_forces_inertial=W_f, joint_force_references=tau)
    data = H.to_product(model, od, td, cuda_device, velocity_representation=js.common.VelRepr.Body)
    t = lambda a: torch.as_tensor(a, dtype=td, device=cuda_device)  # noqa: E731
    out = js.model.step(model, data, link_forces=t(fb), joint_force_references=t(tau))
    H.compare_data(out, ref, H.RTOL[dtype], f"weld step {dtype}")


def test_welded_four_bar_stays_closed(cuda_device):
    """The reference's behavioural check (tests/test_simulations.py:549-612): 1 s of simulation of the welded linkage,
    dropped from 10 cm onto the ground with the joints driven; the two frames coincide at the end."""
    import torch

    case = dict(C.case("four_bar_weld"), constraints=[("tip_a_frame", "tip_b_frame", 1e4, None)])
    model = H.build_model_for_case(case)
    B = 8
    base_position = torch.tensor([0.0, 0.0, 0.10], dtype=torch.float64, device=cuda_device).repeat(B, 1)
    data = js.data.JaxSimModelData.build(model, base_position=base_position, batch_size=B, dtype=torch.float64, device=cuda_device,
                                         velocity_representation=js.common.VelRepr.Inertial)
    tau = torch.linspace(-0.5, 0.5, B, dtype=torch.float64, device=cuda_device)[:, None] * torch.tensor(
        [1.0, 1.0, 0.0, 0.0], dtype=torch.float64, device=cuda_device)
    data = js.model.step_n(model, data, 1000, joint_force_references=tau)
    c = model.kin_dyn_parameters.constraints
    H1 = js.frame.transform(model, data, frame_index=c.frame_idxs_1[0])
    H2 = js.frame.transform(model, data, frame_index=c.frame_idxs_2[0])
    assert torch.isfinite(H1).all() and torch.isfinite(H2).all()
    assert float((H1[:, :3, 3] - H2[:, :3, 3]).abs().max()) < 1e-3
    R_err = H1[:, :3, :3].transpose(-1, -2) @ H2[:, :3, :3]
    assert float((R_err - torch.eye(3, dtype=torch.float64, device=cuda_device)).abs().max()) < 1e-2
    # the unconstrained linkage opens under the same drive: the constraint is what holds it
    free = H.build_model_for_case(dict(case, constraints=None))
    d2 = js.data.JaxSimModelData.build(free, base_position=base_position, batch_size=B, dtype=torch.float64, device=cuda_device,
                                       velocity_representation=js.common.VelRepr.Inertial)
    d2 = js.model.step_n(free, d2, 1000, joint_force_references=tau)
    i1, i2 = c.frame_idxs_1[0], c.frame_idxs_2[0]
    gap = (js.frame.transform(free, d2, frame_index=i1)[:, :3, 3] - js.frame.transform(free, d2, frame_index=i2)[:, :3, 3]).abs().max()
    assert float(gap) > 1e-2


def test_constraints_unsupported_configurations(cuda_device):
    import torch

    from jaxsim_b200.rbda.contacts import RigidContacts

    case = C.case("four_bar_weld")
    model = H.build_model_for_case(dict(case, contact="rigid"))
    assert isinstance(model.contact_model, RigidContacts)
    data = js.data.random_model_data(model, batch_size=2, dtype=torch.float64, device=cuda_device)
    with pytest.raises(NotImplementedError):
        js.model.step(model, data)
    rk = H.build_model_for_case(dict(case, integrator="rk4"))
    with pytest.raises(NotImplementedError):
        js.model.step(rk, js.data.random_model_data(rk, batch_size=2, dtype=torch.float64, device=cuda_device))
