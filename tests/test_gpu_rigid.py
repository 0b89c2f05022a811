"""-m gpu: the rigid-contact step (BASELINE config 3) through the C ABI
against ``oracle/rigid_oracle.py`` on the same inputs.

Parity is defined on the optimum of the contact QP (the reference stops ``qpax`` at
``solver_tol=1e-3``; see the oracle's module docstring), tolerance = north_star's
1e-5 rel (fp64) / 1e-3 rel (fp32) on every leaf of the returned data."""

import dataclasses

import numpy as np
import pytest

import jaxsim_b200.api as js
from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams
from oracle import jaxsim_oracle as O
from oracle import rigid_oracle as R

from . import helpers as H

pytestmark = pytest.mark.gpu


def _dtype(name):
    import torch

    return {"float64": torch.float64, "float32": torch.float32}[name]


def _model(name, **params):
    return H.build_model(name, contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build(**params))


def _f32_representable(od):
    """Round every leaf to float32 and recompute the caches in float64: the oracle then works
    in float64 on exactly the numbers the float32 kernel receives."""
    c = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)  # noqa: E731
    return c


def _inputs(om, B, seed, mode, dtype):
    od = O.random_model_data(om, B, seed=seed, in_contact=mode)
    if dtype == "float32":
        c = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)  # noqa: E731
        od = O.data_replace(om, c(od.joint_positions), c(od.joint_velocities), c(od.base_quaternion),
                            c(od.base_linear_velocity), c(od.base_angular_velocity), c(od.base_position))
    return od


def _vel_floors(od):
    """Scale of the velocities BEFORE the step: the impact can bring them to exactly zero."""
    v = max(float(np.abs(od.base_linear_velocity).max()), float(np.abs(od.base_angular_velocity).max()),
            float(np.abs(od.joint_velocities).max()) if od.joint_velocities.size else 0.0, 1e-3)
    return {"base_linear_velocity": v, "base_angular_velocity": v, "joint_velocities": v, "link_velocities": v}


CASES = [("box", 16), ("sphere", 8), ("icub_like", 8), ("ergocub_like", 4)]


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("mode", [False, True, "flat"])
@pytest.mark.parametrize("name,B", CASES)
def test_rigid_step(name, B, mode, dtype, cuda_device):
    import torch

    model = _model(name, K=1e4, D=20.0)
    om = H.oracle_model(model)
    od = _inputs(om, B, 11, mode, dtype)
    rng = np.random.default_rng(1)
    tau = 10 * rng.uniform(size=(B, om.dofs())).astype(np.float32).astype(np.float64)
    ref = R.step(om, od, joint_force_references=tau)
    pd = H.to_product(model, od, _dtype(dtype), cuda_device)
    out = js.model.step(model, pd, joint_force_references=torch.as_tensor(tau, dtype=_dtype(dtype), device=cuda_device))
    assert not out.contact_state
    H.compare_data(out, ref, H.RTOL[dtype], f"rigid step {name} {mode} {dtype}", floors=_vel_floors(od))


@pytest.mark.parametrize("name,B", [("box", 8), ("icub_like", 4)])
def test_rigid_step_f32_qp_in_f32(name, B, cuda_device):
    """Option B200SIM_OPT_RIGID_QP_F32 (all-float32 contact solve): forces good to ~1e-3 like
    the reference's own solver tolerance, so the step is compared at 2e-2."""
    import torch

    model = _model(name, K=1e4, D=20.0)
    om = H.oracle_model(model)
    od = _inputs(om, B, 5, "flat", "float32")
    ref = R.step(om, od)
    pd = H.to_product(model, od, torch.float32, cuda_device)
    model.set_options(rigid_qp_f32=True)
    out = js.model.step(model, pd)
    H.compare_data(out, ref, 2e-2, f"rigid step {name} qp32", floors=_vel_floors(od))


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_rigid_rollout_uses_stale_link_velocities(dtype, cuda_device):
    """20 consecutive steps: the caches returned by step k (pre-impact link velocities,
    rigid.py:429-434) feed step k+1 exactly like in the reference."""
    name, B = "icub_like", 4
    model = _model(name, K=1e4, D=20.0)
    om = H.oracle_model(model)
    od = _inputs(om, B, 3, "flat", dtype)
    floors = _vel_floors(od)
    pd = H.to_product(model, od, _dtype(dtype), cuda_device)
    for k in range(20):
        od = R.step(om, od)
        pd = js.model.step(model, pd)
    tol = {"float64": 1e-5, "float32": 5e-3}[dtype]
    H.compare_data(pd, od, tol, f"rigid rollout {dtype}", floors=floors)
    # the same 20 steps as ONE step_n call (20 cascades on the stream, stepping in place from the second step on,
    # each reading the caches the previous one wrote): bit-identical to the loop above
    import torch

    fused = js.model.step_n(model, H.to_product(model, _inputs(om, B, 3, "flat", dtype), _dtype(dtype), cuda_device), 20)
    for _, leaf in H.LEAVES:
        assert torch.equal(getattr(fused, leaf), getattr(pd, leaf)), leaf
    # the quirk is observable: the cached link velocities differ from those of the state
    fresh = O.data_replace(om, od.joint_positions, od.joint_velocities, od.base_quaternion, od.base_linear_velocity,
                           od.base_angular_velocity, od.base_position)
    assert np.abs(fresh.link_velocities - od.link_velocities).max() > 1e-6


def test_rigid_box_comes_to_rest(cuda_device):
    """The reference's known answer for this path (tests/test_simulations.py:245-292): a box
    dropped from 2h with K=1e5 and its four bottom corners enabled rests at h/2 after 1 s."""
    import torch

    model = _model("box", K=1e5)
    cp = model.kin_dyn_parameters.contact_parameters
    model.kin_dyn_parameters.contact_parameters = dataclasses.replace(cp, enabled=tuple([True] * 4 + [False] * 4))
    data = js.data.JaxSimModelData.build(model, base_position=torch.tensor([0.0, 0.0, 0.2], dtype=torch.float64),
                                         velocity_representation=js.common.VelRepr.Inertial, batch_size=3,
                                         dtype=torch.float64, device=cuda_device)
    for _ in range(1000):
        data = js.model.step(model, data)
    p = data.base_position.cpu().numpy()
    np.testing.assert_allclose(p[:, 0:2], 0.0, atol=1e-9)
    np.testing.assert_allclose(p[:, 2], 0.05, rtol=1e-7)


def test_rigid_unsupported_configurations(cuda_device):
    import torch

    model = _model("box")
    cp = model.kin_dyn_parameters.contact_parameters
    model.kin_dyn_parameters.contact_parameters = dataclasses.replace(cp, enabled=tuple([False, True] + [True] * 6))
    with pytest.raises(Exception):  # enabled points must be a prefix (rigid.py:401-409 double indexing)
        js.data.JaxSimModelData.build(model, batch_size=2, dtype=torch.float64, device=cuda_device)
    model = _model("box")
    data = js.data.JaxSimModelData.build(model, batch_size=2, dtype=torch.float64, device=cuda_device)
    with pytest.raises(NotImplementedError):  # every rigid step reads the caches the previous one wrote
        js.model.step_n(model, data, 2, update_caches=False)
    js.model.step_n(model, data, 2)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_rigid_cascade_mixed_batch(dtype, cuda_device):
    """Airborne and in-contact environments in one batch: the airborne ones are finished by the
    fused step kernel (level 0 of the cascade), the others by the rigid kernel; an airborne
    environment that lands during the step takes the impact-only route."""
    import torch

    name, B = "icub_like", 24
    model = _model(name, K=1e4, D=20.0)
    om = H.oracle_model(model)
    a = _inputs(om, B // 3, 21, False, dtype)
    b = _inputs(om, B // 3, 22, "flat", dtype)
    c = _inputs(om, B // 3, 23, "flat", dtype)
    # c: just above the ground and falling -> in the air at t, in contact at t+dt
    W_p_Cc, _ = O.collidable_points_pos_vel(om, c.link_transforms, c.link_velocities)
    pc = c.base_position.copy()
    pc[:, 2] += 0.0005 - W_p_Cc[..., 2].min(axis=1)  # lowest point 0.5 mm above the ground, falling at 1 m/s
    vc = c.base_linear_velocity.copy()
    vc[:, 2] = -1.0 - np.cross(c.base_angular_velocity, pc)[:, 2]
    cat = lambda f: np.concatenate([getattr(a, f), getattr(b, f), getattr(c, f)], axis=0)  # noqa: E731
    p = np.concatenate([a.base_position, b.base_position, pc], axis=0)
    v = np.concatenate([a.base_linear_velocity, b.base_linear_velocity, vc], axis=0)
    if dtype == "float32":
        p, v = p.astype(np.float32).astype(np.float64), v.astype(np.float32).astype(np.float64)
    od = O.data_replace(om, cat("joint_positions"), cat("joint_velocities"), cat("base_quaternion"), v,
                        cat("base_angular_velocity"), p)
    W_p_C, _ = O.collidable_points_pos_vel(om, od.link_transforms, od.link_velocities)
    touching = (W_p_C[..., 2] < 0).any(axis=1)
    assert touching[B // 3:2 * B // 3].all() and not touching[2 * B // 3:].any()
    ref = R.step(om, od)
    W_p_C2, _ = O.collidable_points_pos_vel(om, ref.link_transforms, ref.link_velocities)
    assert (W_p_C2[2 * B // 3:, :, 2] < 0).any(axis=1).all()  # group c has landed
    pd = H.to_product(model, od, _dtype(dtype), cuda_device)
    out = js.model.step(model, pd)
    H.compare_data(out, ref, H.RTOL[dtype], f"rigid cascade {dtype}", floors=_vel_floors(od))
    # in place: the result overwrites the input buffers
    pd2 = H.to_product(model, od, _dtype(dtype), cuda_device)
    pd2 = js.model.step(model, pd2, out=pd2)
    for leaf in ("_joint_positions", "_joint_velocities", "_base_linear_velocity", "_link_velocities"):
        assert torch.equal(getattr(pd2, leaf), getattr(out, leaf)), leaf


def test_rigid_more_active_points_than_the_fast_workspace(cuda_device):
    """ErgoCub-like with the terrain above the robot: all 32 collidable points are active,
    which exceeds the 16-point workspace of cascade level 1 -> level 2 (full-size)."""
    import torch

    from jaxsim_b200.terrain import FlatTerrain

    model = H.build_model("ergocub_like", contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build(),
                          terrain=FlatTerrain.build(height=3.0))
    om = H.oracle_model(model)
    od = _inputs(om, 2, 31, "flat", "float64")
    p = od.base_position.copy()
    p[:, 2] -= 2.0
    od = O.data_replace(om, od.joint_positions, od.joint_velocities, od.base_quaternion, od.base_linear_velocity,
                        od.base_angular_velocity, p)
    W_p_C, _ = O.collidable_points_pos_vel(om, od.link_transforms, od.link_velocities)
    assert (W_p_C[..., 2] < 3.0).all()
    ref = R.step(om, od)
    pd = H.to_product(model, od, torch.float64, cuda_device)
    out = js.model.step(model, pd)
    H.compare_data(out, ref, 1e-5, "rigid 32 active points", floors=_vel_floors(od))


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_rigid_split_cascade_equals_monolithic_and_is_graph_capturable(dtype, cuda_device):
    """The default cascade runs level 1 as assemble / solve / resume launches with level 2 on a side stream
    (csrc/b200sim.cu: launch_rigid_split); ``rigid_mono`` keeps the contact QP inside the rigid kernel.  Both call the
    same solver on the same numbers: every leaf agrees bit for bit, on a batch that takes every route of the cascade
    (airborne, few active points, more than the 12 of level 1, landing during the step).  The split cascade is also
    captured in a CUDA graph (the side stream forks and joins inside the capture) and replayed."""
    import torch

    from jaxsim_b200.terrain import FlatTerrain

    B = 96
    om0 = H.oracle_model(_model("ergocub_like", K=1e4, D=20.0))
    a = _inputs(om0, B // 4, 41, False, dtype)     # airborne
    b = _inputs(om0, B // 4, 42, "flat", dtype)    # standing: several active points
    c = _inputs(om0, B // 4, 43, True, dtype)      # touching with a random orientation
    d = _inputs(om0, B // 4, 44, "flat", dtype)    # sunk 4 mm: 14-16 active points -> level 2
    pd_ = d.base_position.copy()
    pd_[:, 2] -= 0.004
    cat = lambda f: np.concatenate([getattr(a, f), getattr(b, f), getattr(c, f), getattr(d, f)], axis=0)  # noqa: E731
    p = np.concatenate([a.base_position, b.base_position, c.base_position, pd_], axis=0)
    perm = np.random.default_rng(5).permutation(B)
    od = O.data_replace(om0, cat("joint_positions")[perm], cat("joint_velocities")[perm], cat("base_quaternion")[perm],
                        cat("base_linear_velocity")[perm], cat("base_angular_velocity")[perm], p[perm])
    W_p_C, _ = O.collidable_points_pos_vel(om0, od.link_transforms, od.link_velocities)
    n_act = (W_p_C[..., 2] < 0).sum(axis=1)
    assert (n_act == 0).any() and ((n_act > 0) & (n_act <= 12)).any() and (n_act > 12).any(), n_act
    outs = {}
    for mono in (False, True):
        model = _model("ergocub_like", K=1e4, D=20.0)
        if mono:
            model.set_options(rigid_mono=True)
        pd = H.to_product(model, od, _dtype(dtype), cuda_device)
        status = torch.zeros(B, dtype=torch.int32, device=cuda_device)
        out = js.model.step(model, pd, status_flags=status)
        out = js.model.step(model, out, status_flags=status)  # second step: stale cached velocities, impacts
        outs[mono] = (model, pd, out, status.clone())
    for leaf in ("_joint_positions", "_joint_velocities", "_base_quaternion", "_base_position", "_base_linear_velocity",
                 "_base_angular_velocity", "_link_transforms", "_link_velocities", "_joint_transforms"):
        x, y = getattr(outs[False][2], leaf), getattr(outs[True][2], leaf)
        assert torch.isfinite(x).all(), leaf
        assert torch.equal(x, y), (leaf, float((x - y).abs().max()))
    assert torch.equal(outs[False][3], outs[True][3])
    # CUDA graph of the split cascade on a side stream of the caller
    model, pd, _, _ = outs[False]
    side = torch.cuda.Stream(cuda_device)
    with torch.cuda.stream(side):
        eager = js.model.step(model, pd)            # scratch of (model, side stream) grows outside the capture
        buf = js.model.step(model, pd)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            js.model.step(model, pd, out=buf)
        for leaf in ("_joint_velocities", "_base_position", "_link_velocities"):
            getattr(buf, leaf).zero_()
        graph.replay()
        side.synchronize()
    for leaf in ("_joint_positions", "_joint_velocities", "_base_position", "_base_linear_velocity", "_link_velocities"):
        assert torch.equal(getattr(buf, leaf), getattr(eager, leaf)), leaf


# ------------------------------------------------------------------------------------------
# RelaxedRigidContacts (rbda/contacts/relaxed_rigid.py): same assembly as the rigid model, the
# contact forces are the solution of (Delassus + diag(r)) x = -b on the active points, no impact
# ------------------------------------------------------------------------------------------
def _relaxed_model(name, **params):
    from jaxsim_b200.rbda.contacts import RelaxedRigidContacts, RelaxedRigidContactsParams

    return H.build_model(name, contact_model=RelaxedRigidContacts.build(), contact_params=RelaxedRigidContactsParams.build(**params))


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("mode", [False, True, "flat"])
@pytest.mark.parametrize("name,B", [("box", 16), ("icub_like", 8), ("ergocub_like", 4)])
def test_relaxed_rigid_step(name, B, mode, dtype, cuda_device):
    import torch

    model = _relaxed_model(name, mu=0.5)
    om = H.oracle_model(model)
    od = _inputs(om, B, 17, mode, dtype)
    rng = np.random.default_rng(2)
    tau = 10 * rng.uniform(size=(B, om.dofs())).astype(np.float32).astype(np.float64)
    ref = R.step(om, od, joint_force_references=tau)
    pd = H.to_product(model, od, _dtype(dtype), cuda_device)
    out = js.model.step(model, pd, joint_force_references=torch.as_tensor(tau, dtype=_dtype(dtype), device=cuda_device))
    assert not out.contact_state
    H.compare_data(out, ref, H.RTOL[dtype], f"relaxed step {name} {mode} {dtype}", floors=_vel_floors(od))


def test_relaxed_rigid_box_settles(cuda_device):
    """A box dropped from 5 mm onto the ground under the relaxed-rigid model comes to rest on it
    (the reference's own scenario for this model, tests/test_simulations.py:295-344): bounded
    penetration, vanishing velocity."""
    import torch

    model = _relaxed_model("box", mu=0.5)
    B = 8
    p = torch.zeros(B, 3, dtype=torch.float64, device=cuda_device)
    p[:, 2] = 0.05 + 0.005
    data = js.data.JaxSimModelData.build(model, base_position=p, batch_size=B, dtype=torch.float64, device=cuda_device,
                                         velocity_representation=js.common.VelRepr.Inertial)
    for _ in range(600):
        data = js.model.step(model, data)
    z = data.base_position[:, 2].cpu().numpy()
    v = data._base_linear_velocity.abs().max().item()
    assert np.all(z < 0.05 + 1e-4) and np.all(z > 0.05 - 2e-3), z
    assert v < 5e-3, v


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_rigid_full_batch_parity(dtype, cuda_device):
    """BASELINE configs[2] at its FULL size: ErgoCub-like, RigidContacts, 16 384 environments, one third airborne,
    one third touching the ground with a random attitude, one third standing flat (many active points) -- the
    work-list cascade at scale (fused kernel / 12-point rigid kernel / full-size rigid kernel).  The NumPy oracle
    steps a strided sample of 258 environments; the whole batch is checked through size-independent properties:
    finite results, unit quaternions, and no enabled collidable point moving INTO the ground after the impact."""
    import torch

    name, B, S = "ergocub_like", 16384, 258
    model = _model(name, K=1e4, D=20.0)
    om = H.oracle_model(model)
    third = B // 3
    parts = [_inputs(om, third, 41, False, dtype), _inputs(om, third, 42, True, dtype), _inputs(om, B - 2 * third, 43, "flat", dtype)]
    cat = lambda f: np.concatenate([getattr(p, f) for p in parts], axis=0)  # noqa: E731
    perm = np.random.default_rng(3).permutation(B)  # interleave the three kinds over the blocks
    od = O.data_replace(om, cat("joint_positions")[perm], cat("joint_velocities")[perm], cat("base_quaternion")[perm],
                        cat("base_linear_velocity")[perm], cat("base_angular_velocity")[perm], cat("base_position")[perm])
    rng = np.random.default_rng(2)
    tau = 10 * rng.uniform(size=(B, om.dofs())).astype(np.float32).astype(np.float64)
    td = _dtype(dtype)
    pd = H.to_product(model, od, td, cuda_device)
    out = js.model.step(model, pd, joint_force_references=torch.as_tensor(tau, dtype=td, device=cuda_device))
    torch.cuda.synchronize()
    # ---- strided sample against the oracle
    idx = np.arange(0, B, B // S)[:S]
    sub = O.data_replace(om, od.joint_positions[idx], od.joint_velocities[idx], od.base_quaternion[idx],
                         od.base_linear_velocity[idx], od.base_angular_velocity[idx], od.base_position[idx])
    ref = R.step(om, sub, joint_force_references=tau[idx])
    it = torch.as_tensor(idx, device=cuda_device)
    sample = js.data._map_leaves(out, lambda t: t[it])
    H.compare_data(sample, ref, H.RTOL[dtype], f"rigid full batch {dtype}", floors=_vel_floors(sub))
    # ---- whole batch
    for _, leaf in H.LEAVES:
        assert bool(torch.isfinite(getattr(out, leaf)).all()), leaf
    qn = torch.linalg.norm(out.base_quaternion, dim=-1)
    assert float((qn - 1).abs().max()) <= (1e-12 if dtype == "float64" else 1e-6)
    # (fresh kinematics: the rigid step returns the PRE-impact link velocities in its cache, rigid.py:429-434)
    fresh = js.data.JaxSimModelData.build(
        model, base_position=out._base_position, base_quaternion=out._base_quaternion, joint_positions=out._joint_positions,
        joint_velocities=out._joint_velocities, base_linear_velocity=out._base_linear_velocity,
        base_angular_velocity=out._base_angular_velocity, velocity_representation=js.common.VelRepr.Inertial,
        batch_size=B, dtype=td, device=cuda_device)
    pos = js.contact.collidable_point_positions(model, fresh)
    vel = js.contact.collidable_point_velocities(model, fresh)
    below = pos[..., 2] < 0
    if bool(below.any()):  # rigid.py:385-436: active points have no velocity into the ground after the impact
        assert float(vel[..., 2][below].min()) >= -(1e-6 if dtype == "float64" else 2e-3)
