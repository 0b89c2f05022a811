"""-m gpu: the CUDA path (through the C ABI, via the Python mirror of the reference API)
against fixtures produced by RUNNING THE REFERENCE's own sources
(tests/golden/make_goldens.py, `oracle/refshim/README.md`).  Same inputs, the reference's
outputs; north_star tolerances 1e-5 rel (float64) / 1e-3 rel (float32) on every leaf.
Nothing here needs `/root/reference`: only the committed .npz fixtures travel to the GPU box.
"""

import json

import numpy as np
import pytest

import jaxsim_b200.api as js
from oracle import jaxsim_oracle as O
from tests.golden import cases as C

from . import helpers as H

pytestmark = pytest.mark.gpu

IDS = [c["id"] for c in C.all_cases()]
VELREPR = {"inertial": js.common.VelRepr.Inertial, "mixed": js.common.VelRepr.Mixed, "body": js.common.VelRepr.Body}


def _dtype(name):
    import torch

    return {"float64": torch.float64, "float32": torch.float32}[name]


def _load(cid):
    z = np.load(C.fixture_path(cid), allow_pickle=False)
    return z, json.loads(str(z["spec"]))


class _Ref:
    """The reference's outputs of a fixture, shaped like the oracle's data for compare_data."""

    def __init__(self, z, soft):
        for oname, pname in H.LEAVES:
            setattr(self, oname, z["out" + pname])
        self.tangential_deformation = z["out_tangential_deformation"] if soft else None


def _inputs(case, z, model, dtype, dev):
    om = H.oracle_model(model)
    od = O.data_replace(om, z["in_joint_positions"], z["in_joint_velocities"], z["in_base_quaternion"],
                        z["in_base_linear_velocity"], z["in_base_angular_velocity"], z["in_base_position"],
                        z["in_tangential_deformation"])
    return H.to_product(model, od, dtype, dev, velocity_representation=VELREPR[case["velrepr"]])


def _vel_floors(z):
    v = max(float(np.abs(z["in_base_linear_velocity"]).max()), float(np.abs(z["in_base_angular_velocity"]).max()),
            float(np.abs(z["in_joint_velocities"]).max()) if z["in_joint_velocities"].size else 0.0, 1e-3)
    return {"base_linear_velocity": v, "base_angular_velocity": v, "joint_velocities": v, "link_velocities": v}


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("cid", IDS)
def test_step_matches_reference(cid, dtype, cuda_device):
    import torch

    z, _ = _load(cid)
    case = C.case(cid)
    if dtype == "float32" and not case["fp32"]:
        pytest.skip("float64-only fixture (tests/golden/cases.py says why)")
    model = H.build_model_for_case(case)
    td = _dtype(dtype)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=td, device=cuda_device)  # noqa: E731
    data = _inputs(case, z, model, td, cuda_device)
    out = data
    for _ in range(case["rollout"]):
        out = js.model.step(model, out, link_forces=t(z["in_link_forces"]) if case["fext"] else None,
                            joint_force_references=t(z["in_tau"]) if case["tau"] else None)
    assert out.velocity_representation == data.velocity_representation
    soft = case["contact"] == "soft"
    # weld constraints: the solve returns wrench pairs of 1e4-1e5 N that cancel to a few N on the mechanism (see
    # tests/golden/cases.py), so the velocity error is set by eps * |wrench| and is measured against the velocity scale of
    # the whole state, not of the (scaled-down) joint rates alone
    floors = _vel_floors(z) if (case["contact"] in ("rigid", "relaxed") or case["rollout"] > 1 or case["constraints"]) else None
    # a rollout accumulates the rounding of its steps (contacts amplify it): 5x the one-step tolerance
    rtol = H.RTOL[dtype] * (5 if case["rollout"] > 1 else 1)
    H.compare_data(out, _Ref(z, soft), rtol, f"golden {cid} {dtype}", floors=floors)
    if case["rollout"] > 1 and soft and not case["fext"]:
        # the fused multi-step launch returns the same state as repeated steps
        outn = js.model.step_n(model, data, case["rollout"], joint_force_references=t(z["in_tau"]) if case["tau"] else None)
        H.compare_data(outn, _Ref(z, soft), rtol, f"golden step_n {cid} {dtype}", floors=floors)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("cid", [c["id"] for c in C.all_cases() if c["rbda"]])
def test_rbda_matches_reference(cid, dtype, cuda_device):
    import torch

    z, _ = _load(cid)
    case = C.case(cid)
    model = H.build_model_for_case(case)
    td = _dtype(dtype)
    rtol = H.RTOL[dtype]
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=td, device=cuda_device)  # noqa: E731
    data = _inputs(case, z, model, td, cuda_device)
    g = lambda x: x.detach().cpu().numpy()  # noqa: E731
    lf = t(z["in_link_forces"]) if case["fext"] else None
    for name, vr in VELREPR.items():
        data.velocity_representation = vr
        vd, sdd = js.model.forward_dynamics_aba(model, data, joint_forces=t(z["in_tau"]), link_forces=lf)
        assert H.rel_err(g(vd), z[f"aba_base_acceleration_{name}"]) <= rtol, name
        assert H.rel_err(g(sdd), z[f"aba_joint_accelerations_{name}"]) <= rtol, name
        fb, tj = js.model.inverse_dynamics(model, data, joint_accelerations=t(z["in_joint_accelerations"]),
                                           base_acceleration=t(z["in_base_acceleration"]), link_forces=lf)
        assert H.rel_err(g(fb), z[f"rnea_base_force_{name}"]) <= rtol, name
        assert H.rel_err(g(tj), z[f"rnea_joint_forces_{name}"]) <= rtol, name
        M = g(js.model.free_floating_mass_matrix(model, data))
        ref = z[f"mass_matrix_{name}"]
        if not model.floating_base():
            M, ref = M[..., -model.dofs():, -model.dofs():], ref[..., 6:, 6:]
        assert H.rel_err(M, ref) <= rtol, name
        assert H.rel_err(g(js.model.free_floating_bias_forces(model, data)), z[f"bias_forces_{name}"]) <= rtol, name
        assert H.rel_err(g(js.model.free_floating_gravity_forces(model, data)), z[f"gravity_forces_{name}"]) <= rtol, name
        if model.floating_base() or name == "body":
            # (fixed base: the reference inverts the full free-floating matrix; compared where the base block
            # needs no change of representation)
            Mi, refi = g(js.model.free_floating_mass_matrix_inverse(model, data)), z[f"mass_matrix_inverse_{name}"]
            assert H.rel_err(Mi, refi) <= (10 * rtol if dtype == "float32" else rtol), name  # inverse: conditioning


@pytest.mark.parametrize("cid", IDS)
def test_contact_helpers_match_reference(cid, cuda_device):
    """`js.contact.collidable_point_kinematics`, `in_contact` and `estimate_good_contact_parameters`
    against the reference's (api/contact.py:18-45, 83-143, 155-211)."""
    import torch

    z, _ = _load(cid)
    case = C.case(cid)
    model = H.build_model_for_case(case)
    data = _inputs(case, z, model, torch.float64, cuda_device)
    pos, vel = js.contact.collidable_point_kinematics(model, data)
    if z["cp_position"].size:
        assert H.rel_err(pos.cpu().numpy(), z["cp_position"]) <= 1e-9
        assert H.rel_err(vel.cpu().numpy(), z["cp_velocity"]) <= 1e-9
        assert np.array_equal(js.contact.in_contact(model, data).cpu().numpy(), z["links_in_contact"] > 0.5)
    for tag, kw in (("default", {}), ("tuned", dict(number_of_active_collidable_points_steady_state=4,
                                                     static_friction_coefficient=1.0, damping_ratio=1.0, max_penetration=0.001))):
        prm = js.contact.estimate_good_contact_parameters(model, device=cuda_device, **kw)
        ref = z[f"egcp_{tag}"][0]
        np.testing.assert_allclose([prm.K, prm.D, prm.mu], ref, rtol=1e-9)


@pytest.mark.parametrize("cid", [c["id"] for c in C.all_cases() if c["rbda"]])
def test_data_build_representations_match_reference(cid, cuda_device):
    """`JaxSimModelData.build` with base velocities given in Inertial / Mixed / Body representation
    (api/data.py:66-202) and the `base_velocity` accessor (api/data.py:288-312)."""
    import torch

    z, _ = _load(cid)
    case = C.case(cid)
    model = H.build_model_for_case(case)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=cuda_device)  # noqa: E731
    B = z["in_base_position"].shape[0]
    data = _inputs(case, z, model, torch.float64, cuda_device)
    for name, vr in VELREPR.items():
        d = js.data.JaxSimModelData.build(
            model, base_position=t(z["in_base_position"]), base_quaternion=t(z["in_base_quaternion"]),
            joint_positions=t(z["in_joint_positions"]), joint_velocities=t(z["in_joint_velocities"]),
            base_linear_velocity=t(z["in_base_linear_velocity"]), base_angular_velocity=t(z["in_base_angular_velocity"]),
            velocity_representation=vr, batch_size=B, dtype=torch.float64, device=cuda_device)
        stored = torch.cat([d._base_linear_velocity, d._base_angular_velocity], dim=-1).cpu().numpy()
        assert H.rel_err(stored, z[f"build_{name}_stored_velocity"]) <= 1e-9, name
        assert H.rel_err(d.link_velocities.cpu().numpy(), z[f"build_{name}_link_velocities"]) <= 1e-9, name
        data.velocity_representation = vr
        assert H.rel_err(data.base_velocity.cpu().numpy(), z[f"base_velocity_{name}"]) <= 1e-9, name
