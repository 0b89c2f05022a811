"""CPU tests of the host logic: URDF loader, model container, C-ABI surface."""

import ctypes
import os
import pathlib
import re

import numpy as np
import pytest

import jaxsim_b200.api as js
from jaxsim_b200 import _lib, models
from jaxsim_b200.api.kin_dyn_parameters import LinkParameters
from jaxsim_b200.parsers.urdf import build_kin_dyn_parameters

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_icub_like_topology():
    name, kd, floating = build_kin_dyn_parameters(models.urdf("icub_like"))
    assert floating and kd.number_of_links() == 24 and kd.number_of_joints() == 23
    assert kd.link_names[0] == "root_link"
    # BFS with children sorted by name (parsers/kinematic_graph.py:669-709)
    assert kd.link_names[1:4] == ("l_hip_1", "r_hip_1", "torso_1")
    assert np.all(kd.parent_array[1:] < np.arange(1, 24)) and kd.parent_array[0] == -1
    # joint index == child link index
    assert kd.joint_model.joint_names[1] == "l_hip_pitch" and kd.joint_model.joint_names[3] == "torso_pitch"
    # levels are contiguous index ranges
    flat = [i for lvl in kd.levels() for i in lvl]
    assert flat == list(range(24))
    # the head (fixed joint) was lumped into the chest and became a frame
    chest = kd.link_names.index("chest")
    assert abs(kd.link_parameters.mass[chest] - 7.8) < 1e-12
    assert "head" in kd.frame_parameters.name and "head" not in kd.link_names
    # two feet x 8 box corners
    assert len(kd.contact_parameters.body) == 16
    assert set(kd.contact_parameters.body) == {kd.link_names.index("l_foot"), kd.link_names.index("r_foot")}
    # URDF convention: suc_H_i = I, lam_H_pre = joint origin
    assert np.allclose(kd.joint_model.suc_H_i, np.eye(4))
    assert np.allclose(kd.motion_subspaces[0], 0)


def test_fixed_base_and_lumping():
    name, kd, floating = build_kin_dyn_parameters(models.urdf("pendulum"))
    assert not floating and kd.link_names == ("base", "arm")
    # world joint origin lands in suc_H_i[0] (math/joint_model.py:78-83, parser.py:192-197)
    assert np.allclose(kd.joint_model.suc_H_i[0][0:3, 3], [0, 0, 1.0])
    assert kd.joint_model.joint_dofs[0] == 0
    # bob (1 kg at -0.5) lumped into the arm (0.5 kg at -0.25): M += X^T M X
    assert abs(kd.link_parameters.mass[1] - 1.5) < 1e-12
    assert np.allclose(kd.link_parameters.center_of_mass[1], [0, 0, -(0.5 * 0.25 + 1.0 * 0.5) / 1.5])
    # collidable points: base box 8 + bob sphere 50 (moved to the arm)
    assert kd.contact_parameters.body.count(0) == 8 and kd.contact_parameters.body.count(1) == 50
    # continuous joint: unbounded limits
    assert kd.joint_parameters.position_limits_max[0] > 1e300


def test_spatial_inertia_roundtrip():
    rng = np.random.default_rng(0)
    _, kd, _ = build_kin_dyn_parameters(models.urdf("ergocub_like"))
    M = kd.link_parameters.spatial_inertias()
    lp2 = LinkParameters.from_spatial_inertias(M)
    assert np.allclose(lp2.mass, kd.link_parameters.mass)
    assert np.allclose(lp2.center_of_mass, kd.link_parameters.center_of_mass)
    assert np.allclose(lp2.inertia_elements, kd.link_parameters.inertia_elements)
    for Mi in M:
        assert np.allclose(Mi, Mi.T) and np.all(np.linalg.eigvalsh(Mi) > 0)


def test_collision_env_vars(monkeypatch):
    monkeypatch.setenv("JAXSIM_COLLISION_USE_BOTTOM_ONLY", "1")
    _, kd, _ = build_kin_dyn_parameters(models.urdf("box"))
    assert len(kd.contact_parameters.body) == 4
    assert np.all(kd.contact_parameters.point[:, 2] < 0)
    monkeypatch.setenv("JAXSIM_COLLISION_USE_BOTTOM_ONLY", "0")
    monkeypatch.setenv("JAXSIM_COLLISION_SPHERE_POINTS", "20")
    _, kd, _ = build_kin_dyn_parameters(models.urdf("sphere"))
    assert len(kd.contact_parameters.body) == 20
    assert np.allclose(np.linalg.norm(kd.contact_parameters.point, axis=1), 0.1)


def test_model_defaults():
    m = js.model.JaxSimModel.build_from_model_description(models.urdf("icub_like"))
    assert m.time_step == 0.001 and m.gravity == -9.81  # api/model.py:54-62,206
    assert type(m.contact_model).__name__ == "SoftContacts"
    assert (m.contact_params.K, m.contact_params.D, m.contact_params.mu) == (1e6, 2000.0, 0.5)
    assert m.actuation_params.torque_max == 3000.0 and m.actuation_params.enable_friction
    assert m.dofs() == 23 and m.floating_base() and len(m.joint_names()) == 23
    rigid = __import__("jaxsim_b200").rbda.contacts.RigidContacts.build()
    mr = js.model.JaxSimModel.build_from_model_description(models.urdf("box"), contact_model=rigid)
    assert type(mr.contact_params).__name__ == "RigidContactsParams" and mr.contact_params.K == 0.0


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "b200sim.h").read_text()
    declared = set(re.findall(r"\b(b200sim_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    lib = _lib.load()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert b"sm_100a" in lib.b200sim_version()
    # descriptor layout: ctypes mirror == C struct (8 int32 + 17 pointers + 19 doubles)
    assert ctypes.sizeof(_lib.B200SimModelDesc) == 8 * 4 + 17 * 8 + 19 * 8


def test_invalid_arguments_are_rejected_without_a_gpu():
    lib = _lib.load()
    out = ctypes.c_void_p()
    assert lib.b200sim_model_create(None, 0, ctypes.byref(out)) == -1
    d = _lib.B200SimModelDesc(abi_version=999)
    assert lib.b200sim_model_create(ctypes.byref(d), 0, ctypes.byref(out)) == -1
    assert lib.b200sim_model_set_tuning(None, 8, 0) == -1
    assert lib.b200sim_step(None, 0, 1, *([None] * 21)) == -1
    assert lib.b200sim_fk(None, 0, 1, *([None] * 12)) == -1
    assert lib.b200sim_aba(None, 0, 1, *([None] * 11)) == -1
    assert lib.b200sim_rnea(None, 0, 1, *([None] * 12)) == -1
    assert lib.b200sim_crba(None, 0, 1, None, None, None) == -1
    assert lib.b200sim_step_n(None, 0, 1, 1, *([None] * 8), 0, None, 0, *([None] * 14)) == -1


def test_product_path_has_no_cpu_fallback():
    import torch

    m = js.model.JaxSimModel.build_from_model_description(models.urdf("box"))
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        js.data.JaxSimModelData.build(m, device="cpu")


def test_product_does_not_import_the_oracle():
    for p in (ROOT / "jaxsim_b200").rglob("*.py"):
        assert "oracle" not in p.read_text(), p


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly ONE line on
    stdout, valid JSON, with the keys of the contract."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--batch", "256"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_relaxed_rigid_model_classes():
    """RelaxedRigidContacts / Params mirror the reference's defaults (rbda/contacts/relaxed_rigid.py:30-82,186-194)."""
    from jaxsim_b200.rbda.contacts import RelaxedRigidContacts, RelaxedRigidContactsParams

    p = RelaxedRigidContactsParams.build()
    assert (p.time_constant, p.damping_coefficient, p.d_min, p.d_max, p.width, p.midpoint, p.power, p.mu) == \
        (0.02, 1.0, 0.9, 0.95, 0.001, 0.5, 2.0, 0.005)
    assert p.valid() and not RelaxedRigidContactsParams.build(d_min=0.99, d_max=0.5).valid()
    m = js.model.JaxSimModel.build_from_model_description(models.urdf("box"), contact_model=RelaxedRigidContacts.build(solver_options={"maxiter": 10}))
    assert isinstance(m.contact_params, RelaxedRigidContactsParams) and dict(m.contact_model.solver_options)["maxiter"] == 10


def test_model_edits_invalidate_the_device_blob():
    """The device blob bakes in time step, gravity, contact / actuation parameters: assigning any of them, or
    `replace`, must drop the cached device models (the reference edits immutable pytrees with replace())."""
    import dataclasses

    from jaxsim_b200.rbda.contacts import SoftContactsParams

    model = js.model.JaxSimModel.build_from_model_description(models.urdf("icub_like"))
    model._devices[0] = object()  # stands for a created device model
    other = dataclasses.replace(model, time_step=2e-3)
    assert other._devices == {} and other._devices is not model._devices and other.time_step == 2e-3
    assert model.replace(gravity=-1.0)._devices == {}
    assert 0 in model._devices  # untouched by the copies
    model._tuning = (8, 0)      # not a physics field
    assert 0 in model._devices
    model.contact_params = SoftContactsParams.build(K=1e5)
    assert model._devices == {}
    model._devices[0] = object()
    model.time_step = 5e-4
    assert model._devices == {}


def test_contact_params_must_match_the_contact_model():
    from jaxsim_b200.rbda.contacts import RigidContacts, SoftContactsParams

    with pytest.raises(TypeError):
        js.model.JaxSimModel.build_from_model_description(
            models.urdf("icub_like"), contact_model=RigidContacts.build(), contact_params=SoftContactsParams.build())


def test_constraint_map_and_frames():
    """ConstraintMap / js.frame host logic (api/kin_dyn_parameters.py:1258-1350, api/frame.py:21-109)."""
    import math

    import jaxsim_b200.api as js
    from jaxsim_b200 import models

    model = js.model.JaxSimModel.build_from_model_description(models.urdf("four_bar"))
    nL = model.number_of_links()
    assert model.kin_dyn_parameters.frame_parameters.name == ("tip_a_frame", "tip_b_frame")
    ia = js.frame.name_to_idx(model, frame_name="tip_a_frame")
    ib = js.frame.name_to_idx(model, frame_name="tip_b_frame")
    assert (ia, ib) == (nL, nL + 1)
    assert js.frame.idx_to_name(model, frame_index=ib) == "tip_b_frame"
    assert model.link_names()[js.frame.idx_of_parent_link(model, frame_index=ia)] == "coupler_a"
    with pytest.raises(ValueError):
        js.frame.name_to_idx(model, frame_name="nope")
    with pytest.raises(ValueError):
        js.frame.idx_of_parent_link(model, frame_index=nL - 1)  # a link index is not a frame index
    empty = js.kin_dyn_parameters.ConstraintMap()
    cmap = empty.add_constraint(model, ia, ib, js.kin_dyn_parameters.ConstraintType.Weld)
    assert len(empty) == 0 and len(cmap) == 1  # immutable: add_constraint returns a new map
    assert cmap.K_P == (1000.0,) and cmap.K_D == (2 * math.sqrt(1000.0),)
    assert model.link_names()[cmap.parent_link_idxs_2[0]] == "coupler_b"
    cmap2 = cmap.add_constraint(model, ib, ia, js.kin_dyn_parameters.ConstraintType.Weld, K_P=1e4, K_D=3.0)
    assert cmap2.K_P == (1000.0, 1e4) and cmap2.K_D[1] == 3.0 and cmap2.frame_idxs_1 == (ia, ib)
    with pytest.raises(NotImplementedError):
        cmap.add_constraint(model, ia, ib, 1)
    with model.editable(validate=False) as edited:
        edited.kin_dyn_parameters.constraints = cmap
    assert model.kin_dyn_parameters.constraints is None and len(edited.kin_dyn_parameters.constraints) == 1


def test_sdf_pose_semantics_and_errors():
    """parsers/sdf.py: poses resolve through `relative_to` chains with the SDF 1.7+ default frames (link -> model,
    joint -> child link, frame -> attached_to); what the loader cannot give the reference's meaning to raises."""
    from jaxsim_b200 import models
    from jaxsim_b200.parsers.urdf import build_kin_dyn_parameters

    name, kd, floating = build_kin_dyn_parameters(models.urdf("posed_sdf"))
    assert name == "posed" and floating and kd.link_names[0] == "trunk" and kd.number_of_links() == 4
    jm = kd.joint_model
    # base link posed in the model frame -> suc_H_i[0]; the hip is given in the MODEL frame, so lam_H_pre = trunk^-1 hip
    M_H_trunk = jm.suc_H_i[0]
    assert np.allclose(M_H_trunk[0:3, 3], [0.1, 0.2, 0.3])
    i_thigh = kd.link_names.index("thigh")
    M_H_hip = M_H_trunk @ jm.lam_H_pre[i_thigh]
    assert np.allclose(M_H_hip[0:3, 3], [0.1, 0.3, 0.3], atol=1e-12)
    assert np.allclose(jm.suc_H_i[i_thigh], np.eye(4), atol=1e-12)          # thigh posed at its joint
    i_shin = kd.link_names.index("shin")
    assert not np.allclose(jm.suc_H_i[i_shin], np.eye(4))                    # knee posed 5 cm above its child link
    assert np.allclose(np.linalg.inv(jm.suc_H_i[i_shin])[0:3, 3], [0, 0, 0.05], atol=1e-12)
    assert kd.frame_parameters.name == ("camera",) and int(kd.frame_parameters.body[0]) == 0
    k = kd.joint_model.joint_names.index("hip") - 1
    assert kd.joint_parameters.position_limit_spring[k] == 50.0 and kd.joint_parameters.position_limit_damper[k] == 2.0

    def sdf(body):
        return f'<?xml version="1.0"?><sdf version="1.9"><model name="m">{body}</model></sdf>'

    link = ('<link name="{n}"><pose relative_to="{r}">0 0 0 0 0 0</pose><inertial><mass>1</mass><inertia><ixx>1</ixx><iyy>1</iyy>'
            '<izz>1</izz></inertia></inertial></link>')
    with pytest.raises(ValueError, match="cyclic"):
        build_kin_dyn_parameters(sdf(link.format(n="a", r="b") + link.format(n="b", r="a")))
    with pytest.raises(ValueError, match="unknown frame"):
        build_kin_dyn_parameters(sdf(link.format(n="a", r="nowhere")))
    with pytest.raises(ValueError, match="not supported"):
        build_kin_dyn_parameters(sdf(link.format(n="a", r="__model__") + link.format(n="b", r="a")
                                     + '<joint name="j" type="ball"><parent>a</parent><child>b</child></joint>'))
    with pytest.raises(NotImplementedError, match="inertial"):
        build_kin_dyn_parameters(sdf('<link name="a"><inertial><pose relative_to="__model__">0 0 0 0 0 0</pose><mass>1</mass></inertial></link>'))


def test_option_flags_match_the_header():
    """The Python mirror of the implementation switches and include/b200sim.h agree (incl. the round-2 rigid switch)."""
    import pathlib
    import re

    from jaxsim_b200 import _lib

    hdr = (pathlib.Path(__file__).resolve().parent.parent / "include" / "b200sim.h").read_text()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define B200SIM_OPT_(\w+) (\d+)", hdr)}
    assert flags["RIGID_MONO"] == 128
    for name, value in flags.items():
        assert getattr(_lib, "OPT_" + name) == value, name
