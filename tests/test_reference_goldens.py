"""-m "not gpu": the oracle and the URDF loader against fixtures produced by RUNNING THE
REFERENCE's own sources (tests/golden/make_goldens.py; `oracle/refshim/README.md` explains
how the reference is executed without JAX).  This is what pins the oracle: every fixture
holds the inputs and what `jaxsim.api.model.step` (and `forward_dynamics_aba`,
`inverse_dynamics`, `free_floating_mass_matrix`, `ode.system_dynamics`) returned for them.

Tolerances: 1e-9 relative for everything that is closed-form (soft contacts, RK4, RBDAs:
two float64 evaluation orders of the same formulas); 1e-6 for the rigid-contact step, whose
contact forces are the optimum of a QP found by two different interior-point codes.
"""

import json

import numpy as np
import pytest

from oracle import jaxsim_oracle as O
from oracle import constraints_oracle as KC
from oracle import rigid_oracle as R
from tests.golden import cases as C

from . import helpers as H

IDS = [c["id"] for c in C.all_cases()]
OUT_LEAVES = [(o, "out" + p) for o, p in H.LEAVES]


def _load(cid):
    path = C.fixture_path(cid)
    if not path.exists():
        pytest.fail(f"golden fixture {path.name} is missing: run tests/golden/make_goldens.py in the build container")
    z = np.load(path, allow_pickle=False)
    spec = json.loads(str(z["spec"]))
    return z, spec


def _models(case):
    from jaxsim_b200 import models

    pm = H.build_model_for_case(case)
    return pm, H.oracle_model(pm), models.urdf(case["model"])


def _oracle_data(om, z):
    return O.data_replace(om, z["in_joint_positions"], z["in_joint_velocities"], z["in_base_quaternion"],
                          z["in_base_linear_velocity"], z["in_base_angular_velocity"], z["in_base_position"],
                          z["in_tangential_deformation"])


def _link_forces_inertial(case, z, od):
    if not case["fext"]:
        return None
    return O.other_representation_to_inertial(z["in_link_forces"], case["velrepr"], od.link_transforms, is_force=True)


def _rel(a, b, floor=1e-12):
    return float(np.max(np.abs(np.asarray(a, float) - np.asarray(b, float)))) / max(float(np.max(np.abs(b))), floor) if np.size(b) else 0.0


@pytest.mark.parametrize("cid", IDS)
def test_fixture_matches_case_table(cid):
    """The fixture was generated for the case as it is defined today, on today's URDF text."""
    z, spec = _load(cid)
    case = C.case(cid)
    _, _, urdf = _models(case)
    assert spec["urdf_sha256_16"] == C.urdf_digest(urdf), "model URDF changed since the fixture was generated"
    for k, v in case.items():
        if k not in spec:  # an option added after this fixture was generated: it must be at its default
            assert v == C.DEFAULTS[k], (k, v)
            continue
        assert spec[k] == v or (isinstance(v, (dict, list)) and json.loads(json.dumps(v)) == spec[k]), (k, v, spec[k])


@pytest.mark.parametrize("cid", sorted({c["id"] for c in C.all_cases() if c["id"] in
                                        ("pendulum_soft", "double_pendulum_soft", "cartpole_soft", "box_soft_air", "sphere_soft_contact",
                                         "icub_soft_air", "ergocub_soft_contact")}))
def test_urdf_loader_matches_reference_kinematic_graph(cid):
    """`jaxsim_b200.parsers.urdf` against the reference's KinDynParameters.build +
    ModelDescription.build_model_from (fixed-joint lumping, BFS link order, joint model,
    collidable points)."""
    z, _ = _load(cid)
    pm, _, _ = _models(C.case(cid))
    kd = pm.kin_dyn_parameters
    assert tuple(kd.link_names) == tuple(str(s) for s in z["kd_link_names"])
    assert np.array_equal(np.asarray(kd.parent_array), z["kd_parent_array"])
    assert bool(pm.floating_base()) == bool(z["kd_floating_base"])
    lp, jm, cp, jp = kd.link_parameters, kd.joint_model, kd.contact_parameters, kd.joint_parameters
    np.testing.assert_allclose(lp.mass, z["kd_mass"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(lp.center_of_mass, z["kd_center_of_mass"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(lp.inertia_elements, z["kd_inertia_elements"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(jm.lam_H_pre, z["kd_lam_H_pre"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(jm.suc_H_i, z["kd_suc_H_i"], rtol=1e-12, atol=1e-14)
    assert tuple(int(t) for t in jm.joint_types) == tuple(int(t) for t in z["kd_joint_types"])
    np.testing.assert_allclose(np.asarray(jm.joint_axis).reshape(-1, 3), z["kd_joint_axis"].reshape(-1, 3), rtol=1e-12, atol=1e-14)
    assert tuple(int(b) for b in cp.body) == tuple(int(b) for b in z["kd_contact_body"])
    np.testing.assert_allclose(np.asarray(cp.point).reshape(-1, 3), z["kd_contact_point"].reshape(-1, 3), rtol=1e-12, atol=1e-14)
    assert tuple(bool(b) for b in cp.enabled) == tuple(bool(b) for b in z["kd_contact_enabled"])
    for name in ("friction_static", "friction_viscous", "position_limits_min", "position_limits_max",
                 "position_limit_spring", "position_limit_damper"):
        np.testing.assert_allclose(getattr(jp, name), z["kd_" + name], rtol=1e-12, atol=0, err_msg=name)


@pytest.mark.parametrize("cid", IDS)
def test_oracle_step_matches_reference(cid):
    z, _ = _load(cid)
    case = C.case(cid)
    _, om, _ = _models(case)
    od = _oracle_data(om, z)
    # the caches the reference built for the INPUT state are what the oracle's data_replace gives
    W_f = _link_forces_inertial(case, z, od)
    tau = z["in_tau"] if case["tau"] else None
    out = od
    for _ in range(case["rollout"]):
        if case["constraints"]:
            out, tol = KC.step(om, out, link_forces_inertial=W_f, joint_force_references=tau), 1e-9
        elif case["contact"] in ("rigid", "relaxed"):
            out, tol = R.step(om, out, link_forces_inertial=W_f, joint_force_references=tau), 1e-6
        elif case["integrator"] == "rk4fast":
            out, tol = O.step_rk4fast(om, out, link_forces_inertial=W_f, joint_force_references=tau), 1e-9
        elif case["integrator"] == "rk4":
            out, tol = O.step_rk4(om, out, link_forces_inertial=W_f, joint_force_references=tau), 1e-9
        else:
            out, tol = O.step(om, out, link_forces_inertial=W_f, joint_force_references=tau), 1e-9
    if case["rollout"] > 1:
        tol *= 10
    # velocities can be brought to ~0 by a rigid impact: measure them against their scale before the step
    vscale = max(float(np.abs(z["in_base_linear_velocity"]).max()), float(np.abs(z["in_base_angular_velocity"]).max()),
                 float(np.abs(z["in_joint_velocities"]).max()) if z["in_joint_velocities"].size else 0.0, 1e-3)
    errs = {}
    for oname, key in OUT_LEAVES:
        floor = vscale if ("velocit" in oname and case["contact"] in ("rigid", "relaxed")) else 1e-12
        errs[oname] = _rel(getattr(out, oname), z[key], floor)
    if case["contact"] == "soft":
        errs["tangential_deformation"] = _rel(out.tangential_deformation, z["out_tangential_deformation"], 1e-6)
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f"{cid}: oracle differs from the reference beyond {tol}: {bad}"


@pytest.mark.parametrize("cid", [c["id"] for c in C.all_cases() if c["rbda"]])
def test_oracle_rbda_matches_reference(cid):
    z, _ = _load(cid)
    case = C.case(cid)
    _, om, _ = _models(case)
    od = _oracle_data(om, z)
    B, nL = od.joint_positions.shape[0], om.number_of_links()
    # the "_inertial" goldens read the link forces in inertial-fixed representation
    W_f = z["in_link_forces"] if case["fext"] else np.zeros((B, nL, 6))
    args = (od.base_position, od.base_orientation, od.joint_positions, od.base_linear_velocity, od.base_angular_velocity,
            od.joint_velocities)
    vd, sdd = O.aba(om, *args, z["in_tau"], W_f)
    assert _rel(vd, z["aba_base_acceleration_inertial"]) <= 1e-9
    assert _rel(sdd, z["aba_joint_accelerations_inertial"]) <= 1e-9
    fb, tj = O.rnea(om, *args, z["in_base_acceleration"], z["in_joint_accelerations"], W_f)
    assert _rel(fb, z["rnea_base_force_inertial"]) <= 1e-9
    assert _rel(tj, z["rnea_joint_forces_inertial"]) <= 1e-9
    # joint accelerations do not depend on the representation the base quantities are expressed in,
    # provided the link forces are the same physical forces
    if not case["fext"]:
        for name in ("mixed", "body"):
            assert _rel(sdd, z[f"aba_joint_accelerations_{name}"]) <= 1e-9
    M = O.crba(om, od.joint_positions)
    if om.floating_base:
        assert _rel(M, z["mass_matrix_body"]) <= 1e-9
    else:
        assert _rel(M[:, 6:, 6:], z["mass_matrix_body"][:, 6:, 6:]) <= 1e-9  # joint block of a fixed-base model
    if "ode_joint_velocities" in z.files:
        xd = O.system_dynamics(om, od, np.zeros((B, nL, 6)), z["in_tau"])  # joint torques as given (api/ode.py:174-225)
        for k in ("base_position", "base_quaternion", "joint_positions", "base_linear_velocity", "base_angular_velocity",
                  "joint_velocities", "tangential_deformation"):
            assert _rel(xd[k], z["ode_" + k], 1e-9) <= 1e-9, k


@pytest.mark.parametrize("cid", IDS)
def test_oracle_collidable_point_kinematics_match_reference(cid):
    z, _ = _load(cid)
    case = C.case(cid)
    pm, om, _ = _models(case)
    od = _oracle_data(om, z)
    W_p, W_pd = O.collidable_points_pos_vel(om, od.link_transforms, od.link_velocities)
    en = [k for k, e in enumerate(pm.kin_dyn_parameters.contact_parameters.enabled) if e]
    assert _rel(W_p[:, en], z["cp_position"]) <= 1e-9
    assert _rel(W_pd[:, en], z["cp_velocity"], 1e-9) <= 1e-9


BRANCHED_URDF = """<robot name="branched">
  <link name="base"><inertial><origin xyz="0.01 0 0.02" rpy="0.1 0.2 0.3"/><mass value="2"/><inertia ixx="0.02" iyy="0.03" izz="0.04" ixy="0.001" ixz="0" iyz="0.002"/></inertial>
    <collision><origin xyz="0 0 -0.05" rpy="0 0 0.3"/><geometry><box size="0.2 0.1 0.05"/></geometry></collision></link>
  <link name="z_sensor"><inertial><origin xyz="0 0.01 0"/><mass value="0.2"/><inertia ixx="0.001" iyy="0.001" izz="0.001"/></inertial>
    <collision><origin xyz="0.02 0 0"/><geometry><sphere radius="0.03"/></geometry></collision></link>
  <joint name="sensor_fix" type="fixed"><origin xyz="0.1 0 0.05" rpy="0 0.5 0"/><parent link="base"/><child link="z_sensor"/></joint>
  <link name="b_arm"><inertial><origin xyz="0 0 0.1"/><mass value="1"/><inertia ixx="0.01" iyy="0.01" izz="0.002"/></inertial>
    <collision><origin xyz="0 0 0.2"/><geometry><box size="0.04 0.04 0.04"/></geometry></collision></link>
  <joint name="j_b" type="revolute"><origin xyz="0 0.1 0" rpy="0.2 0 0"/><parent link="base"/><child link="b_arm"/><axis xyz="0 1 0"/><limit lower="-1" upper="1" effort="10" velocity="5"/><dynamics damping="0.1" friction="0.05"/></joint>
  <link name="b_tip"><inertial><origin xyz="0 0 0.02"/><mass value="0.1"/><inertia ixx="0.0001" iyy="0.0001" izz="0.0001"/></inertial>
    <collision><origin xyz="0 0 0.03"/><geometry><sphere radius="0.02"/></geometry></collision></link>
  <joint name="tip_fix" type="fixed"><origin xyz="0 0 0.25" rpy="0 0 1.0"/><parent link="b_arm"/><child link="b_tip"/></joint>
  <link name="a_arm"><inertial><origin xyz="0.05 0 0"/><mass value="0.8"/><inertia ixx="0.002" iyy="0.008" izz="0.008"/></inertial>
    <collision><origin xyz="0.1 0 0"/><geometry><box size="0.1 0.03 0.03"/></geometry></collision></link>
  <joint name="j_a" type="prismatic"><origin xyz="0 -0.1 0"/><parent link="base"/><child link="a_arm"/><axis xyz="1 0 0"/><limit lower="-0.2" upper="0.3" effort="10" velocity="5"/></joint>
</robot>"""


# model files the reference ships (read in place: never copied into this repository; build container only); the last one
# is SDF: poses `relative_to` other frames, resolved by parsers/sdf.py and, on the reference's side, by the stand-in's
# `switch_frame_convention` (tests/conftest.py:718 loads it the same way)
REFERENCE_URDFS = {"ref:cartpole": "examples/assets/cartpole.urdf", "ref:4_bar_opened": "tests/assets/4_bar_opened.urdf",
                   "ref:double_pendulum.sdf": "tests/assets/double_pendulum.sdf"}

# SDF pose semantics beyond the reference's asset: a floating base whose link frame is posed in the model frame, a joint
# posed in the model frame (not in its child), a child link offset from its joint (non-identity successor transform), an
# explicit <frame> attached to a link, <limit> stiffness / dissipation
@pytest.mark.parametrize("name", ["pendulum", "double_pendulum", "cartpole", "box", "sphere", "icub_like", "ergocub_like", "four_bar",
                                  "four_bar_fixed", "branched", "posed_sdf", *REFERENCE_URDFS])
def test_urdf_front_end_matches_reference_parser(name):
    """The product's URDF loader against the reference's OWN front end run on the same URDF text:
    `jaxsim.parsers.rod.build_model_description` = parsers/rod/parser.py:36-420 + parsers/rod/utils.py:21-225 (inertial ->
    6D inertia, box / sphere -> collidable points, joint limits / friction, env-var knobs) + the kinematic-graph code,
    executed over the `rod` stand-in of oracle/refshim (which only builds rod's dataclass tree from the URDF elements).
    Build container only: skipped where /root/reference is absent."""
    from tests.golden import refenv

    if not (refenv.REFERENCE_SRC / "jaxsim").is_dir():
        pytest.skip("reference sources not available")
    from jaxsim_b200 import models
    from jaxsim_b200.parsers.urdf import build_kin_dyn_parameters

    jaxsim, js = refenv.load()
    # "branched": rotated inertial frames, fixed-joint lumping of links WITH collision shapes on two branches (the
    # reference keeps the shapes in file order and only re-parents them: ADVICE r1), prismatic + revolute joints
    if name in REFERENCE_URDFS:  # the reference's own model files: its example cart-pole and the 4-bar linkage of its tests
        path = refenv.REFERENCE_SRC.parent / REFERENCE_URDFS[name]
        if not path.is_file():
            pytest.skip(f"{path} not available")
        text = path.read_text()
    else:
        text = BRANCHED_URDF if name == "branched" else models.urdf(name)  # "posed_sdf" is an SDF document (jaxsim_b200/models)
    ref = js.model.JaxSimModel.build(model_description=refenv.reference_model_description(text), time_step=1e-3,
                                     gravity=-jaxsim.math.STANDARD_GRAVITY)
    rk = ref.kin_dyn_parameters
    _, kd, floating = build_kin_dyn_parameters(text)
    assert tuple(kd.link_names) == tuple(rk.link_names) and bool(floating) == bool(ref.floating_base())
    assert np.array_equal(np.asarray(kd.parent_array), np.asarray(rk.parent_array))
    np.testing.assert_allclose(kd.link_parameters.mass, np.asarray(rk.link_parameters.mass), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(kd.link_parameters.center_of_mass, np.asarray(rk.link_parameters.center_of_mass), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(kd.link_parameters.inertia_elements, np.asarray(rk.link_parameters.inertia_elements), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(kd.joint_model.lam_H_pre, np.asarray(rk.joint_model.λ_H_pre), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(kd.joint_model.suc_H_i, np.asarray(rk.joint_model.suc_H_i), rtol=1e-12, atol=1e-14)
    assert tuple(int(b) for b in kd.contact_parameters.body) == tuple(int(b) for b in rk.contact_parameters.body)
    np.testing.assert_allclose(np.asarray(kd.contact_parameters.point).reshape(-1, 3), np.asarray(rk.contact_parameters.point).reshape(-1, 3),
                               rtol=1e-12, atol=1e-14)
    if kd.number_of_joints() > 0:
        for f in ("friction_static", "friction_viscous", "position_limits_min", "position_limits_max", "position_limit_spring",
                  "position_limit_damper"):
            np.testing.assert_allclose(getattr(kd.joint_parameters, f), np.asarray(getattr(rk.joint_parameters, f)), rtol=1e-12, atol=0, err_msg=f)
    # frames: massless links on fixed joints (parsers/rod/parser.py:90-124, api/kin_dyn_parameters.py:843-917)
    fp, rf = kd.frame_parameters, rk.frame_parameters
    assert tuple(fp.name) == tuple(rf.name) and tuple(int(b) for b in fp.body) == tuple(int(b) for b in np.asarray(rf.body).reshape(-1))
    if len(fp.name):
        np.testing.assert_allclose(np.asarray(fp.transform), np.asarray(rf.transform), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("cid", [c["id"] for c in C.all_cases() if c["constraints"]])
def test_oracle_constraint_wrenches_match_reference(cid):
    """`compute_constraint_wrenches` (rbda/kinematic_constraints.py:172-345) and `js.frame.transform` of the constrained
    frames, as the reference evaluated them on the input state with no other forces."""
    z, _ = _load(cid)
    case = C.case(cid)
    _, om, _ = _models(case)
    od = _oracle_data(om, z)
    B, nL, n = case["B"], om.number_of_links(), om.dofs()
    for b in range(B):
        W = KC.compute_constraint_wrenches(om, od, b, np.zeros((nL, 6)), np.zeros(n))
        assert _rel(W, z["constraint_wrenches_free"][b]) <= 1e-9
        T = KC.constraint_transforms(om, R._Env(od, b))
        assert _rel(T.reshape(-1, 4, 4), z["constraint_frame_transforms"][b]) <= 1e-12
    # equal and opposite forces on the two frames (the torques differ by the lever arm between the frames)
    W = z["constraint_wrenches_free"]
    np.testing.assert_allclose(W[:, :, 0, 0:3], -W[:, :, 1, 0:3], rtol=1e-12, atol=1e-12)


def test_urdf_loader_frames_match_reference():
    """Massless links on fixed joints become frames of their parent link (parsers/rod/parser.py:90-124,
    api/kin_dyn_parameters.py:843-917): names, parent link and L_H_F as the reference's front end produced them."""
    z, _ = _load("four_bar_weld")
    pm, _, _ = _models(C.case("four_bar_weld"))
    fp = pm.kin_dyn_parameters.frame_parameters
    assert tuple(fp.name) == tuple(str(s) for s in z["kd_frame_names"])
    assert tuple(int(b) for b in fp.body) == tuple(int(b) for b in z["kd_frame_body"])
    np.testing.assert_allclose(np.asarray(fp.transform), z["kd_frame_transform"], rtol=1e-12, atol=1e-14)
    c = pm.kin_dyn_parameters.constraints
    assert (c.frame_idxs_1, c.frame_idxs_2) == ((pm.number_of_links(),), (pm.number_of_links() + 1,))
