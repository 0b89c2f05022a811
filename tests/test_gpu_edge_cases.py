"""-m gpu: edge cases of the step path against the oracle -- the code paths the default
models do not exercise (SURVEY.md 8c "edge cases the reference tests")."""

import copy

import numpy as np
import pytest

import jaxsim_b200.api as js
from jaxsim_b200 import models
from jaxsim_b200.rbda.actuation import ActuationParams
from jaxsim_b200.rbda.contacts import SoftContactsParams
from jaxsim_b200.terrain import FlatTerrain
from oracle import jaxsim_oracle as O

from . import helpers as H

pytestmark = pytest.mark.gpu


def _run(model, om, od, cuda_device, tau=None, dtype="float64", rtol=None):
    import torch

    td = {"float64": torch.float64, "float32": torch.float32}[dtype]
    pd = H.to_product(model, od, td, cuda_device)
    t = None if tau is None else torch.as_tensor(tau, dtype=td, device=cuda_device)
    out = js.model.step(model, pd, joint_force_references=t)
    ref = O.step(om, od, joint_force_references=tau)
    H.compare_data(out, ref, rtol if rtol is not None else H.RTOL[dtype], "edge")
    return out, ref


@pytest.mark.parametrize("B", [1, 3, 5, 37, 4097])
def test_ragged_batch_sizes(B, cuda_device):
    """Batches that do not fill a warp / a block / the grid (idle groups shadow the last env)."""
    model = H.build_model("icub_like")
    om = H.oracle_model(model)
    od = O.random_model_data(om, B, seed=B, in_contact=True)
    _run(model, om, od, cuda_device)


def test_empty_batch_and_unbatched(cuda_device):
    import torch

    model = H.build_model("double_pendulum")
    om = H.oracle_model(model)
    # B = 0: nothing to do, shapes preserved
    d0 = js.data.JaxSimModelData.build(model, batch_size=0, dtype=torch.float64, device=cuda_device,
                                       velocity_representation=js.common.VelRepr.Inertial)
    o0 = js.model.step(model, d0)
    assert o0.joint_positions.shape == (0, 2)
    # unbatched leaves (the reference's non-vmapped call)
    od = O.random_model_data(om, 1, seed=4)
    d1 = js.data.JaxSimModelData.build(
        model, joint_positions=torch.as_tensor(od.joint_positions[0]), joint_velocities=torch.as_tensor(od.joint_velocities[0]),
        dtype=torch.float64, device=cuda_device, velocity_representation=js.common.VelRepr.Inertial)
    assert d1.joint_positions.shape == (2,) and d1.link_transforms.shape == (3, 4, 4)
    o1 = js.model.step(model, d1)
    ref = O.step(om, od)
    assert o1.joint_positions.shape == (2,)
    assert H.rel_err(o1.joint_positions.cpu().numpy()[None], ref.joint_positions) <= 1e-9
    assert H.rel_err(o1.link_transforms.cpu().numpy()[None], ref.link_transforms) <= 1e-9


def test_shape_errors_raise_value_error(cuda_device):
    import torch

    model = H.build_model("icub_like")
    d = js.data.random_model_data(model, batch_size=4, dtype=torch.float32, device=cuda_device)
    with pytest.raises(ValueError):
        js.model.step(model, d, joint_force_references=torch.zeros(4, 22, device=cuda_device))
    with pytest.raises(ValueError):
        js.model.step(model, d, link_forces=torch.zeros(4, 23, 6, device=cuda_device))
    with pytest.raises(ValueError):
        js.data.JaxSimModelData.build(model, joint_positions=torch.zeros(4, 5), device=cuda_device)
    with pytest.raises(TypeError):
        js.model.step(model, js.data.random_model_data(model, batch_size=2, dtype=torch.float16, device=cuda_device))


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_non_default_contact_and_actuation_parameters(dtype, cuda_device):
    """pow() exponents != 0.5, friction disabled, active joint limits with spring + damper,
    tight torque-speed curve, raised terrain, weaker gravity, larger time step."""
    kw = dict(
        contact_params=SoftContactsParams.build(K=2e5, D=300.0, mu=0.8, p=0.7, q=0.3),
        actuation_params=ActuationParams(torque_max=8.0, omega_th=0.4, omega_max=0.9, enable_friction=False),
        terrain=FlatTerrain.build(height=0.07), time_step=2e-3, gravity=3.7,
    )
    model = H.build_model("icub_like", **kw)
    B = 24
    # sample inside the URDF limits first, THEN narrow the limits so that many joints sit beyond them
    od = O.random_model_data(H.oracle_model(model), B, seed=71, in_contact=True)
    jp = model.kin_dyn_parameters.joint_parameters
    jp.position_limit_spring = np.full(23, 40.0)
    jp.position_limit_damper = np.full(23, 0.3)
    jp.position_limits_min = np.full(23, -0.2)
    jp.position_limits_max = np.full(23, 0.25)
    om = H.oracle_model(model)
    assert (om.p, om.q, om.gravity, om.terrain_height) == (0.7, 0.3, -3.7, 0.07)
    rng = np.random.default_rng(3)
    tau = 20 * rng.uniform(-1, 1, size=(B, 23))
    od.joint_velocities = 1.5 * rng.uniform(-1, 1, size=(B, 23))
    od = O.data_replace(om, od.joint_positions, od.joint_velocities, od.base_quaternion, od.base_linear_velocity,
                        od.base_angular_velocity, od.base_position)
    od.tangential_deformation = 1e-4 * rng.uniform(-1, 1, size=od.tangential_deformation.shape)
    out, ref = _run(model, om, od, cuda_device, tau=tau, dtype=dtype)
    # the limits and the tn-curve really were active in this test
    assert np.any(od.joint_positions > 0.25) and np.any(np.abs(od.joint_velocities) > 0.9)


def test_disabled_collidable_points(cuda_device):
    """Disabled points produce no force and their deformation does not change
    (rbda/contacts/soft.py:425-442)."""
    model = H.build_model("box")
    enabled = [True, False, True, False, False, True, True, False]
    model.kin_dyn_parameters.contact_parameters.enabled = tuple(enabled)
    om = H.oracle_model(model)
    od = O.random_model_data(om, 16, seed=5, in_contact=True)
    od.tangential_deformation = 1e-4 * np.random.default_rng(0).uniform(-1, 1, size=od.tangential_deformation.shape)
    out, ref = _run(model, om, od, cuda_device)
    m_out = out.contact_state["tangential_deformation"].cpu().numpy()
    off = [k for k, e in enumerate(enabled) if not e]
    assert np.array_equal(m_out[:, off], od.tangential_deformation[:, off])


def test_fixed_base_with_rotated_mount_and_base_velocity(cuda_device):
    """F_GENERIC_FK: suc_H_i[0] != I (rotated + translated world joint) and the reference's
    quirks for fixed-base models (ABA ignores suc_H_i[0] and the base velocity, FK and the
    contacts honour them; rbda/aba.py:79-121 vs rbda/forward_kinematics.py:66-73)."""
    from jaxsim_b200.models import UrdfBuilder, box_inertia, cylinder_inertia, sphere_inertia

    b = UrdfBuilder("tilted_arm")
    b.massless_link("world")
    b.link("base", 2.0, box_inertia(2.0, (0.2, 0.2, 0.1)), collisions=[("box", (0, 0, 0), (0, 0, 0), (0.2, 0.2, 0.1))])
    b.link("l1", 1.0, cylinder_inertia(1.0, 0.03, 0.4), com=(0, 0, 0.2))
    b.link("l2", 0.5, sphere_inertia(0.5, 0.05), com=(0.1, 0, 0), collisions=[("sphere", (0.1, 0, 0), (0, 0, 0), 0.05)])
    b.joint("mount", "fixed", "world", "base", xyz=(0.3, -0.2, 0.06), rpy=(0.4, -0.3, 0.9))
    b.joint("j1", "revolute", "base", "l1", xyz=(0, 0, 0.05), rpy=(0.2, 0, 0), axis=(0, 1, 0), limit=(-2, 2), damping=0.1)
    b.joint("j2", "prismatic", "l1", "l2", xyz=(0, 0, 0.4), axis=(0.6, 0, 0.8), limit=(-0.2, 0.2), friction=0.3)
    model = js.model.JaxSimModel.build_from_model_description(b.urdf())
    assert not model.floating_base()
    om = H.oracle_model(model)
    B = 12
    od = O.random_model_data(om, B, seed=9)
    rng = np.random.default_rng(1)
    # a user may hand over any base pose/velocity even for fixed-base models
    od = O.data_replace(om, od.joint_positions, od.joint_velocities, O._quat_from_euler_xyz_intrinsic(rng.uniform(-0.5, 0.5, (B, 3))),
                        rng.uniform(-0.3, 0.3, (B, 3)), rng.uniform(-0.3, 0.3, (B, 3)), rng.uniform(-0.1, 0.1, (B, 3)))
    tau = rng.uniform(-2, 2, size=(B, 2))
    _run(model, om, od, cuda_device, tau=tau)
    _run(model, om, od, cuda_device, tau=tau, dtype="float32")


def test_non_identity_successor_transforms(cuda_device):
    """F_SUC_NONID: suc_H_i[i>=1] != I (SDF-style link poses, math/joint_model.py:92-98).
    The URDF loader never produces them, so the KinDynParameters are edited directly."""
    model = H.build_model("icub_like")
    kd = copy.deepcopy(model.kin_dyn_parameters)
    rng = np.random.default_rng(5)
    for i in range(1, kd.number_of_links()):
        e = rng.uniform(-0.3, 0.3, 3)
        q = O._quat_from_euler_xyz_intrinsic(e[None])[0]
        H4 = np.eye(4)
        H4[0:3, 0:3] = O.quat_to_dcm(q)
        H4[0:3, 3] = rng.uniform(-0.02, 0.02, 3)
        kd.joint_model.suc_H_i[i] = H4
    model2 = js.model.JaxSimModel.build(kd, floating_base=True, model_name="icub_suc")
    om = H.oracle_model(model2)
    od = O.random_model_data(om, 10, seed=13, in_contact=True)
    _run(model2, om, od, cuda_device, tau=rng.uniform(-3, 3, (10, 23)))


def test_model_loaded_from_sdf(cuda_device):
    """A model description in SDF (jaxsim_b200/parsers/sdf.py: poses `relative_to` other frames -> lam_H_pre, non-identity
    suc_H_i, posed base link): the loader's result is pinned against the reference's front end on the CPU
    (test_urdf_front_end_matches_reference_parser[posed_sdf]); here the kernels step it like the oracle does."""
    model = H.build_model("posed_sdf")
    kd = model.kin_dyn_parameters
    assert not np.allclose(kd.joint_model.suc_H_i[0], np.eye(4)) and not all(np.allclose(h, np.eye(4)) for h in kd.joint_model.suc_H_i[1:])
    om = H.oracle_model(model)
    od = O.random_model_data(om, 12, seed=17, in_contact=True)
    _run(model, om, od, cuda_device, tau=np.random.default_rng(3).uniform(-2, 2, (12, model.dofs())))


@pytest.mark.parametrize("vr", ["Body", "Mixed", "Inertial"])
def test_build_converts_base_velocity_representation(vr, cuda_device):
    """JaxSimModelData.build stores the base velocity inertial-fixed whatever representation it
    is given in (api/data.py:151-156); base_velocity converts back (api/data.py:288-312)."""
    import torch

    model = H.build_model("icub_like")
    om = H.oracle_model(model)
    od = O.random_model_data(om, 6, seed=3)
    repr_ = getattr(js.common.VelRepr, vr)
    v_in = np.random.default_rng(2).uniform(-1, 1, size=(6, 6))
    W_v = O.other_representation_to_inertial(v_in, vr.lower(), O.transform_from_quat_pos(od.base_quaternion, od.base_position), is_force=False)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=cuda_device)  # noqa: E731
    d = js.data.JaxSimModelData.build(
        model, base_position=t(od.base_position), base_quaternion=t(od.base_quaternion), joint_positions=t(od.joint_positions),
        joint_velocities=t(od.joint_velocities), base_linear_velocity=t(v_in[:, 0:3]), base_angular_velocity=t(v_in[:, 3:6]),
        velocity_representation=repr_, dtype=torch.float64, device=cuda_device)
    got = torch.cat([d._base_linear_velocity, d._base_angular_velocity], -1).cpu().numpy()
    assert H.rel_err(got, W_v) <= 1e-12
    assert H.rel_err(d.base_velocity.cpu().numpy(), v_in) <= 1e-12
    ref = O.data_replace(om, od.joint_positions, od.joint_velocities, od.base_quaternion, W_v[:, 0:3], W_v[:, 3:6], od.base_position)
    assert H.rel_err(d.link_velocities.cpu().numpy(), ref.link_velocities) <= 1e-10
    # replace() keeps the representation and accepts velocities in it
    d2 = d.replace(model, base_linear_velocity=t(v_in[:, 0:3]), base_angular_velocity=t(v_in[:, 3:6]))
    assert d2.velocity_representation == repr_
    assert H.rel_err(d2._base_linear_velocity.cpu().numpy(), W_v[:, 0:3]) <= 1e-12


def test_update_link_params_changes_the_dynamics(cuda_device):
    """b200sim_model_update_link_params (HW-parametrisation hook): new masses take effect and
    match an oracle model built with them."""
    import ctypes

    import torch

    from jaxsim_b200 import _lib

    model = H.build_model("icub_like")
    om = H.oracle_model(model)
    od = O.random_model_data(om, 8, seed=21)
    pd = H.to_product(model, od, torch.float64, cuda_device)
    lp = model.kin_dyn_parameters.link_parameters
    new_mass = lp.mass * np.linspace(0.5, 1.5, lp.mass.size)
    om2 = copy.deepcopy(om)
    om2.kin_dyn_parameters.link_parameters.mass = new_mass
    dm = model.device_model(cuda_device)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(_lib.c_dp)  # noqa: E731
    keep = [np.ascontiguousarray(new_mass), np.ascontiguousarray(lp.center_of_mass), np.ascontiguousarray(lp.inertia_elements)]
    rc = _lib.load().b200sim_model_update_link_params(dm.handle, *[k.ctypes.data_as(_lib.c_dp) for k in keep])
    assert rc == 0
    out = js.model.step(model, pd)
    H.compare_data(out, O.step(om2, od), 1e-5, "updated masses")
    assert H.rel_err(out.joint_velocities.cpu().numpy(), O.step(om, od).joint_velocities) > 1e-4


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_cached_input_kinematics_path(dtype, cuda_device):
    """step() reads the link transforms / velocities cached in the input data (like the
    reference's contact code, api/contact.py:39-43) instead of recomputing them; both paths
    agree with the oracle and with each other, over consecutive steps."""
    import torch

    td = {"float64": torch.float64, "float32": torch.float32}[dtype]
    for name in ("icub_like", "ergocub_like"):
        model = H.build_model(name)
        om = H.oracle_model(model)
        od = O.random_model_data(om, 31, seed=91, in_contact=True)
        rng = np.random.default_rng(4)
        tau = 4 * rng.uniform(-1, 1, size=(31, om.dofs()))
        t = torch.as_tensor(tau, dtype=td, device=cuda_device)
        a = H.to_product(model, od, td, cuda_device)
        b = H.to_product(model, od, td, cuda_device)
        ref = od
        for _ in range(3):
            a = js.model.step(model, a, joint_force_references=t)                          # cached inputs
            b = js.model.step(model, b, joint_force_references=t, use_input_caches=False)  # recomputed
            ref = O.step(om, ref, joint_force_references=tau)
        H.compare_data(a, ref, 3 * H.RTOL[dtype], f"cached {name}")
        H.compare_data(b, ref, 3 * H.RTOL[dtype], f"recomputed {name}")
        tol = 1e-10 if dtype == "float64" else 2e-3  # fp32 + stiff contacts amplify rounding over 3 steps
        for _, leaf in H.LEAVES:
            x, y = getattr(a, leaf), getattr(b, leaf)
            assert float((x - y).abs().max()) <= tol * max(float(y.abs().max()), 1e-9), (name, leaf)


@pytest.mark.gpu
def test_model_edits_take_effect_after_the_first_step(cuda_device):
    """time_step / contact_params assigned after the first step (or `replace`) must reach the device."""
    import torch

    from jaxsim_b200.rbda.contacts import SoftContactsParams

    model = H.build_model("icub_like")
    om = H.oracle_model(model)
    od = O.random_model_data(om, 8, seed=3, in_contact=True)
    pd = H.to_product(model, od, torch.float64, cuda_device)
    js.model.step(model, pd)  # creates the device blob with dt = 1e-3
    model.time_step = 2.5e-4
    model.contact_params = SoftContactsParams.build(K=2e5, D=300.0, mu=0.8)
    out = js.model.step(model, pd)
    ref = O.step(H.oracle_model(model), od)
    H.compare_data(out, ref, H.RTOL["float64"], "edited model")
    other = model.replace(time_step=1e-3)
    H.compare_data(js.model.step(other, pd), O.step(H.oracle_model(other), od), H.RTOL["float64"], "replaced model")
    H.compare_data(js.model.step(model, pd), ref, H.RTOL["float64"], "original after replace")


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_step_n_with_moving_frame_link_forces(dtype, cuda_device):
    """Body / Mixed link_forces follow the links.  The kernels re-express them with the link poses of EVERY
    fused step (b200sim_step_n_ex), so step_n equals repeated step -- which equals the oracle fed with the forces
    converted by the reference's formula at every state (api/model.py:2641-2646)."""
    import torch

    td = torch.float64 if dtype == "float64" else torch.float32
    model = H.build_model("icub_like")
    om = H.oracle_model(model)
    B, T = 6, 4
    od = O.random_model_data(om, B, seed=5, in_contact=True)
    rng = np.random.default_rng(8)
    f_np = rng.uniform(-20, 20, size=(B, model.number_of_links(), 6))
    f = torch.as_tensor(f_np, dtype=td, device=cuda_device)
    for name, vr in (("mixed", js.common.VelRepr.Mixed), ("body", js.common.VelRepr.Body)):
        pd = H.to_product(model, od, td, cuda_device, velocity_representation=vr)
        fused = js.model.step_n(model, pd, T, link_forces=f)
        seq, ref = pd, od
        for _ in range(T):
            seq = js.model.step(model, seq, link_forces=f)
            W_f = O.other_representation_to_inertial(f_np, name, ref.link_transforms, is_force=True)
            ref = O.step(om, ref, link_forces_inertial=W_f)
        # (not bit-identical: repeated steps read the cached inertial-fixed link velocities back, v + p x w - p x w)
        H.compare_data(fused, ref, 5 * H.RTOL[dtype], f"step_n link_forces {name} {dtype}")
        H.compare_data(seq, ref, 5 * H.RTOL[dtype], f"repeated step link_forces {name} {dtype}")
    # a model whose base link pose is offset from the chain root keeps the single-step shim and refuses fused steps
    fixed = H.build_model("pendulum")
    omf = H.oracle_model(fixed)
    odf = O.random_model_data(omf, 3, seed=2)
    pdf = H.to_product(fixed, odf, td, cuda_device, velocity_representation=js.common.VelRepr.Mixed)
    ff = torch.ones(3, fixed.number_of_links(), 6, dtype=td, device=cuda_device)
    js.model.step(fixed, pdf, link_forces=ff)
    with pytest.raises(NotImplementedError):
        js.model.step_n(fixed, pdf, 3, link_forces=ff)


@pytest.mark.gpu
def test_rk4_step_accepts_unbatched_data(cuda_device):
    import torch

    model = H.build_model("icub_like", integrator=js.model.IntegratorType.RungeKutta4)
    om = H.oracle_model(model)
    od = O.random_model_data(om, 2, seed=9, in_contact=True)
    pd = H.to_product(model, od, torch.float64, cuda_device)
    full = js.model.step(model, pd)
    one = js.model.step(model, js.data._map_leaves(pd, lambda t: t[0]))
    assert one.joint_positions.dim() == 1
    assert torch.allclose(one.joint_positions, full.joint_positions[0], rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_rk4_step_is_one_abi_call_and_graph_capturable(dtype, cuda_device):
    """b200sim_step_rk4: the whole RungeKutta4 step runs as launches on the caller's stream (no host round trip), so it
    can be captured in a CUDA graph once its scratch exists; replays give the eager result bit for bit, and external
    forces in Body representation plus joint references match the oracle."""
    import torch

    td = torch.float64 if dtype == "float64" else torch.float32
    model = H.build_model("icub_like", integrator=js.model.IntegratorType.RungeKutta4)
    om = H.oracle_model(model)
    B = 33
    od = O.random_model_data(om, B, seed=21, in_contact=True)
    rng = np.random.default_rng(3)
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)  # noqa: E731
    od = O.data_replace(om, f32(od.joint_positions), f32(od.joint_velocities), f32(od.base_quaternion), f32(od.base_linear_velocity),
                        f32(od.base_angular_velocity), f32(od.base_position),
                        f32(1e-4 * rng.uniform(-1, 1, (B, model.number_of_collidable_points(), 3))))
    tau = f32(rng.uniform(-5, 5, (B, om.dofs())))
    fb = f32(rng.uniform(-5, 5, (B, om.number_of_links(), 6)))
    ref = O.step_rk4(om, od, link_forces_inertial=O.other_representation_to_inertial(fb, "body", od.link_transforms, is_force=True),
                     joint_force_references=tau)
    data = H.to_product(model, od, td, cuda_device, velocity_representation=js.common.VelRepr.Body)
    t = lambda a: torch.as_tensor(a, dtype=td, device=cuda_device)  # noqa: E731
    lf, jf = t(fb), t(tau)
    eager = js.model.step(model, data, link_forces=lf, joint_force_references=jf)
    H.compare_data(eager, ref, H.RTOL[dtype], f"rk4 abi {dtype}")
    side = torch.cuda.Stream(device=cuda_device)
    side.wait_stream(torch.cuda.current_stream(cuda_device))
    with torch.cuda.stream(side):
        js.model.step(model, data, link_forces=lf, joint_force_references=jf)  # creates the scratch of this stream
    torch.cuda.current_stream(cuda_device).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        captured = js.model.step(model, data, link_forces=lf, joint_force_references=jf)
    for leaf in ("_joint_positions", "_base_quaternion", "_link_velocities"):
        getattr(captured, leaf).zero_()
    graph.replay()
    torch.cuda.synchronize(cuda_device)
    for _, leaf in H.LEAVES:
        assert torch.equal(getattr(captured, leaf), getattr(eager, leaf)), leaf
    assert torch.equal(captured.contact_state["tangential_deformation"], eager.contact_state["tangential_deformation"])


@pytest.mark.gpu
@pytest.mark.parametrize("contact", ["soft", "rigid"])
def test_status_flags(contact, cuda_device):
    """Device-side status flags (b200sim_step_n_status): what rbda/utils.py:136-146 raises under
    JAXSIM_ENABLE_EXCEPTIONS, per environment."""
    import torch

    from jaxsim_b200 import _lib
    from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams

    kw = dict(contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build()) if contact == "rigid" else {}
    model = H.build_model("icub_like", **kw)
    om = H.oracle_model(model)
    B = 37
    od = O.random_model_data(om, B, seed=13, in_contact="flat")
    pd = H.to_product(model, od, torch.float64, cuda_device)
    pd._base_quaternion[3, 1] = float("nan")
    pd._base_quaternion[5] *= 1.01
    pd._joint_velocities[7, 2] = float("inf")
    flags = torch.full((B,), -1, dtype=torch.int32, device=cuda_device)
    out = js.model.step(model, pd, status_flags=flags, use_input_caches=False)
    f = flags.cpu().numpy()
    assert f[3] & _lib.STATUS_QUATERNION_NAN and f[3] & _lib.STATUS_NON_FINITE
    assert f[5] == _lib.STATUS_QUATERNION_NOT_UNIT
    assert f[7] & _lib.STATUS_NON_FINITE
    clean = np.ones(B, dtype=bool)
    clean[[3, 5, 7]] = False
    assert not (f[clean] & (_lib.STATUS_QUATERNION_NAN | _lib.STATUS_QUATERNION_NOT_UNIT | _lib.STATUS_NON_FINITE)).any()
    assert bool(torch.isfinite(out._joint_positions[clean]).all())
    with pytest.raises(ValueError):
        js.model.step(model, pd, status_flags=flags, out=pd)
