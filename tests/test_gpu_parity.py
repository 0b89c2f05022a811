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


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("G", [8, 16])
def test_step_wide_levels_and_large_trees(G, dtype, cuda_device):
    """ErgoCub-like tree (50 links, a level of 10 links): with 8 lanes a level spans two packed
    rows and the compact final phase does not apply (nL > 4 G + 1); with 16 lanes both do."""
    import torch

    model = H.build_model("ergocub_like")
    model.set_tuning(lanes_per_env=G)
    om = H.oracle_model(model)
    B = 37
    od = O.random_model_data(om, B, seed=21, in_contact=True)
    tau = 5 * np.random.default_rng(4).uniform(-1, 1, size=(B, om.dofs()))
    ref = O.step(om, od, joint_force_references=tau)
    pd = H.to_product(model, od, _dtype(dtype), cuda_device)
    t = torch.as_tensor(tau, dtype=_dtype(dtype), device=cuda_device)
    out = js.model.step(model, pd, joint_force_references=t)
    H.compare_data(out, ref, H.RTOL[dtype], f"ergocub G={G} {dtype}")
    out3 = js.model.step_n(model, pd, 3, joint_force_references=t)
    ref3 = ref
    for _ in range(2):
        ref3 = O.step(om, ref3, joint_force_references=tau)
    H.compare_data(out3, ref3, 5 * H.RTOL[dtype], f"ergocub step_n G={G} {dtype}")


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("options", [dict(no_bulk_in=True), dict(pdl=False), dict(generic_kernel=True), dict(step_v1=True),
                                     dict(step_v1=True, bulk_in=True), dict(step_v1=True, tma_store=False)])
def test_step_implementation_switches_agree(options, dtype, cuda_device):
    """The implementation switches (per-link cp.async instead of TMA bulk input loads, programmatic dependent
    launch off, generic kernel instance, first-generation specialised kernel and its own switches) never change
    results: bit-identical to the default, or equal to rounding where a different kernel instance runs."""
    import torch

    base = H.build_model("icub_like")
    alt = H.build_model("icub_like")
    alt.set_options(**options)
    om = H.oracle_model(base)
    B = 133  # ragged: not a multiple of the environments per block
    od = O.random_model_data(om, B, seed=23, in_contact=True)
    tau = torch.as_tensor(3 * np.random.default_rng(5).uniform(-1, 1, size=(B, om.dofs())), dtype=_dtype(dtype), device=cuda_device)
    a = js.model.step(base, H.to_product(base, od, _dtype(dtype), cuda_device), joint_force_references=tau)
    b = js.model.step(alt, H.to_product(alt, od, _dtype(dtype), cuda_device), joint_force_references=tau)
    # another kernel instance may schedule the same arithmetic differently (FMA contraction, summation order of siblings)
    exact = "generic_kernel" not in options and "step_v1" not in options
    for _, leaf in H.LEAVES:
        x, y = getattr(a, leaf), getattr(b, leaf)
        if exact:
            assert torch.equal(x, y), (options, leaf)
        else:
            assert H.rel_err(y.cpu().numpy(), x.cpu().numpy()) <= H.RTOL[dtype] * 1e-2, (options, leaf)
    for _ in range(3):  # back-to-back launches (PDL overlap) keep giving the same answer
        b = js.model.step(alt, b, joint_force_references=tau)
        a = js.model.step(base, a, joint_force_references=tau)
    assert H.rel_err(b._joint_positions.cpu().numpy(), a._joint_positions.cpu().numpy()) <= (0 if exact else H.RTOL[dtype])


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
    """200 consecutive steps of boxes dropped on the ground stay within tolerance (fp64).
    Contact parameters from the reference's estimate_good_contact_parameters recipe
    (tests/test_simulations.py:206-214): the default K=1e6/D=2e3 is unstable for a 1 kg box
    at dt=1e-3 in the reference as well (the oracle diverges to 1e21)."""
    import torch

    from jaxsim_b200.rbda.contacts import SoftContactsParams

    K = 1.0 * 9.81 / 4 / 1e-3**1.5
    prm = SoftContactsParams.build(K=K, D=2 * np.sqrt(K * 1.0), mu=0.5)
    model = H.build_model("box", contact_params=prm)
    om = H.oracle_model(model)
    B = 6
    od = O.random_model_data(om, B, seed=17, base_pos_bounds=((-1, -1, 0.2), (1, 1, 0.3)))
    pd = H.to_product(model, od, torch.float64, cuda_device)
    for _ in range(200):
        od = O.step(om, od)
        pd = js.model.step(model, pd)
    assert np.all(np.isfinite(od.base_position)) and np.abs(od.base_position).max() < 10
    assert np.abs(od.tangential_deformation).max() > 0  # the boxes did touch the ground
    H.compare_data(pd, od, 1e-5, "box rollout")


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_step_n_equals_repeated_step(dtype, cuda_device):
    """One fused launch of T steps == T single-step launches (same arithmetic), with
    per-step torque rows, and both match the oracle."""
    import torch

    model = H.build_model("icub_like")
    om = H.oracle_model(model)
    B, T = 19, 7
    od = O.random_model_data(om, B, seed=23, in_contact=True)
    rng = np.random.default_rng(5)
    tau = 5 * rng.uniform(-1, 1, size=(T, B, om.dofs()))
    td = _dtype(dtype)
    pd = H.to_product(model, od, td, cuda_device)
    tau_t = torch.as_tensor(tau, dtype=td, device=cuda_device)
    a = pd
    for k in range(T):
        a = js.model.step(model, a, joint_force_references=tau_t[k])
        od = O.step(om, od, joint_force_references=tau[k])
    b = js.model.step_n(model, pd, T, joint_force_references=tau_t)
    # same algorithm, same inputs: equal up to the rounding of differently-contracted FMAs
    # (the kinematics of step k+1 are inlined at another call site than those of a fresh call)
    tol = 1e-11 if dtype == "float64" else 2e-4
    worst = 0.0
    for _, leaf in H.LEAVES:
        x, y = getattr(a, leaf), getattr(b, leaf)
        worst = max(worst, float((x - y).abs().max()) / max(float(x.abs().max()), 1e-12))
    print(f"step_n vs repeated step ({dtype}): max rel diff {worst:.3e}")
    assert worst <= tol, worst
    dm = (a.contact_state["tangential_deformation"] - b.contact_state["tangential_deformation"]).abs().max()
    assert float(dm) <= tol * 1e-2
    H.compare_data(b, od, 10 * H.RTOL[dtype] if dtype == "float32" else H.RTOL[dtype], f"step_n {dtype}")
    # constant references + no caches
    c = js.model.step_n(model, pd, 3, joint_force_references=tau_t[0], update_caches=False)
    d = pd
    for _ in range(3):
        d = js.model.step(model, d, joint_force_references=tau_t[0])
    assert float((c.joint_positions - d.joint_positions).abs().max()) <= tol and c._link_transforms is None
    with pytest.raises(RuntimeError):
        _ = c.link_transforms


def test_store_paths_and_out_buffers_agree(cuda_device):
    """TMA bulk stores vs 128-bit stores give identical bytes; `out=` reuses buffers."""
    import torch

    od = None
    outs = []
    for tma in (True, False):
        model = H.build_model("ergocub_like")
        model.set_options(tma_store=tma)
        om = H.oracle_model(model)
        od = O.random_model_data(om, 77, seed=29, in_contact=True)
        pd = H.to_product(model, od, torch.float32, cuda_device)
        o1 = js.model.step(model, pd)
        o2 = js.model.step(model, pd, out=js.model.step(model, pd))
        for _, leaf in H.LEAVES:
            assert torch.equal(getattr(o1, leaf), getattr(o2, leaf)), leaf
        outs.append(o1)
        H.compare_data(o1, O.step(om, od), 1e-3, f"ergocub tma={tma}")
    for _, leaf in H.LEAVES:
        assert torch.equal(getattr(outs[0], leaf), getattr(outs[1], leaf)), leaf


def test_in_place_step_and_cuda_graph(cuda_device):
    """`out=data` steps in place; a captured graph of steps replays to the same result."""
    import torch

    model = H.build_model("icub_like")
    om = H.oracle_model(model)
    od = O.random_model_data(om, 64, seed=31)
    ref = H.to_product(model, od, torch.float32, cuda_device)
    for _ in range(4):
        ref = js.model.step(model, ref)
    a = H.to_product(model, od, torch.float32, cuda_device)
    b = H.to_product(model, od, torch.float32, cuda_device)
    js.model.step(model, a, out=a)  # warm-up outside capture (lazy init), then rewind
    a = H.to_product(model, od, torch.float32, cuda_device)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            js.model.step(model, a, out=b)
            js.model.step(model, b, out=a)
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    g.replay()
    torch.cuda.synchronize()
    for _, leaf in H.LEAVES:
        assert torch.equal(getattr(a, leaf), getattr(ref, leaf)), leaf


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", MODELS)
def test_rnea(name, dtype, cuda_device):
    """inverse_dynamics == oracle rnea (rbda/rnea.py) and RNEA(ABA(tau)) == tau on the GPU."""
    import torch

    model = H.build_model(name)
    om = H.oracle_model(model)
    B = 29
    od = O.random_model_data(om, B, seed=37)
    rng = np.random.default_rng(3)
    sdd = rng.uniform(-3, 3, size=(B, om.dofs()))
    avd = rng.uniform(-2, 2, size=(B, 6))
    W_f = rng.uniform(size=(B, om.number_of_links(), 6))
    fB_ref, tau_ref = O.rnea(om, od.base_position, od.base_orientation, od.joint_positions, od.base_linear_velocity,
                             od.base_angular_velocity, od.joint_velocities, avd, sdd, W_f)
    td = _dtype(dtype)
    pd = H.to_product(model, od, td, cuda_device)
    t = lambda a: torch.as_tensor(a, dtype=td, device=cuda_device)  # noqa: E731
    fB, tau = js.model.inverse_dynamics(model, pd, joint_accelerations=t(sdd), base_acceleration=t(avd), link_forces=t(W_f))
    rt = H.RTOL[dtype]
    if om.floating_base:
        assert H.rel_err(fB.cpu().numpy(), fB_ref) <= rt
    else:
        assert float(fB.abs().max()) == 0.0
    if om.dofs():
        assert H.rel_err(tau.cpu().numpy(), tau_ref) <= rt
        # FD/ID consistency on the device (tests/test_api_model.py:495-577)
        tau_in = t(10 * rng.uniform(size=(B, om.dofs())))
        W_f2 = W_f.copy()
        if not om.floating_base:
            W_f2[:, 0] = 0
        a, sdd2 = js.model.forward_dynamics_aba(model, pd, joint_forces=tau_in, link_forces=t(W_f2))
        fB2, tau2 = js.model.inverse_dynamics(model, pd, joint_accelerations=sdd2, base_acceleration=a, link_forces=t(W_f2))
        scale = float(tau_in.abs().max())
        assert float((tau2 - tau_in).abs().max()) / scale <= (1e-9 if dtype == "float64" else 2e-3)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", MODELS)
def test_crba(name, dtype, cuda_device):
    """free_floating_mass_matrix == oracle crba (rbda/crba.py), body-fixed representation."""
    model = H.build_model(name)
    om = H.oracle_model(model)
    B = 21
    od = O.random_model_data(om, B, seed=41)
    M_ref = O.crba(om, od.joint_positions)
    pd = H.to_product(model, od, _dtype(dtype), cuda_device, velocity_representation=js.common.VelRepr.Body)
    M = js.model.free_floating_mass_matrix(model, pd)
    assert M.shape == M_ref.shape
    assert H.rel_err(M.cpu().numpy(), M_ref) <= H.RTOL[dtype]
    Mn = M.cpu().numpy().astype(np.float64)
    assert np.abs(Mn - np.swapaxes(Mn, 1, 2)).max() == 0.0  # exactly symmetric by construction


@pytest.mark.parametrize("what", ["joint_positions", "link_masses", "state", "torques"])
@pytest.mark.parametrize("name", ["icub_like", "double_pendulum", "box"])
def test_step_jvp_matches_finite_differences(name, what, cuda_device):
    """BASELINE config 5: d(step)/d(joint q, link masses, ...) as a JVP on the GPU ==
    central finite differences of the fp64 oracle (the reference validates its AD the same
    way, tests/test_automatic_differentiation.py:24-27,346-420)."""
    import copy

    import torch

    model = H.build_model(name)
    om = H.oracle_model(model)
    n, nL = om.dofs(), om.number_of_links()
    if what in ("joint_positions", "torques") and n == 0:
        pytest.skip("no joints")
    B = 11
    od = O.random_model_data(om, B, seed=43, in_contact=(name != "double_pendulum"))
    rng = np.random.default_rng(7)
    tau = 3 * rng.uniform(-1, 1, size=(B, n))
    eps = 1e-6
    tang, dmass = {}, None

    def perturbed(sign):
        d = copy.deepcopy(od)
        m2 = om
        t = tau
        if what == "joint_positions":
            d.joint_positions = od.joint_positions + sign * eps * tang["joint_positions"]
        elif what == "state":
            d.joint_velocities = od.joint_velocities + sign * eps * tang["joint_velocities"]
            d.base_linear_velocity = od.base_linear_velocity + sign * eps * tang["base_linear_velocity"]
            d.base_angular_velocity = od.base_angular_velocity + sign * eps * tang["base_angular_velocity"]
            d.base_position = od.base_position + sign * eps * tang["base_position"]
            d.base_quaternion = od.base_quaternion + sign * eps * tang["base_quaternion"]
        elif what == "torques":
            t = tau + sign * eps * tang["joint_force_references"]
        elif what == "link_masses":
            m2 = copy.deepcopy(om)
            m2.kin_dyn_parameters.link_parameters.mass = om.kin_dyn_parameters.link_parameters.mass + sign * eps * dmass
        d = O.data_replace(m2, d.joint_positions, d.joint_velocities, d.base_quaternion, d.base_linear_velocity,
                           d.base_angular_velocity, d.base_position, d.tangential_deformation)
        # data_replace normalises q; the step normalises it anyway (base_orientation)
        return O.step(m2, d, joint_force_references=t)

    if what == "joint_positions":
        tang["joint_positions"] = rng.uniform(-1, 1, size=(B, n))
    elif what == "state":
        tang["joint_velocities"] = rng.uniform(-1, 1, size=(B, n))
        for k, w in (("base_linear_velocity", 3), ("base_angular_velocity", 3), ("base_position", 3), ("base_quaternion", 4)):
            tang[k] = rng.uniform(-1, 1, size=(B, w)) * (1.0 if om.floating_base else 0.0)
        tang["base_position"][:, 2] = 0.0  # keep the contact set fixed under the perturbation
    elif what == "torques":
        tang["joint_force_references"] = rng.uniform(-1, 1, size=(B, n))
    else:
        dmass = rng.uniform(0.1, 1, size=nL)
        tang["link_masses"] = dmass
    fp, fm = perturbed(+1), perturbed(-1)
    pd = H.to_product(model, od, torch.float64, cuda_device)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=cuda_device)  # noqa: E731
    out, dout = js.model.step_jvp(model, pd, {k: t(v) for k, v in tang.items()}, joint_force_references=t(tau))
    H.compare_data(out, O.step(om, od, joint_force_references=tau), 1e-9, "jvp primal")
    worst = {}
    for oname, pname in H.LEAVES:
        fd = (getattr(fp, oname) - getattr(fm, oname)) / (2 * eps)
        got = getattr(dout, pname).cpu().numpy()
        if fd.size == 0:
            continue
        scale = max(float(np.abs(fd).max()), 1e-3)
        worst[oname] = float(np.abs(got - fd).max()) / scale
    fd = (fp.tangential_deformation - fm.tangential_deformation) / (2 * eps)
    if fd.size:
        got = dout.contact_state["tangential_deformation"].cpu().numpy()
        worst["tangential_deformation"] = float(np.abs(got - fd).max()) / max(float(np.abs(fd).max()), 1e-3)
    bad = {k: v for k, v in worst.items() if not v <= 2e-5}
    assert not bad, (bad, worst)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", ["icub_like", "double_pendulum", "box", "cartpole"])
def test_step_rk4(name, dtype, cuda_device):
    """IntegratorType.RungeKutta4 (api/integrators.py:91-156): 4 system_dynamics launches
    + host-side combination == oracle rk4; also pins b200sim_dynamics against the oracle."""
    import torch

    model = H.build_model(name, integrator=js.model.IntegratorType.RungeKutta4)
    om = H.oracle_model(model)
    B = 17
    od = O.random_model_data(om, B, seed=47, in_contact=(name in ("icub_like", "box")))
    rng = np.random.default_rng(9)
    tau = 5 * rng.uniform(-1, 1, size=(B, om.dofs()))
    td = _dtype(dtype)
    pd = H.to_product(model, od, td, cuda_device)
    t = lambda a: torch.as_tensor(a, dtype=td, device=cuda_device)  # noqa: E731
    # the state derivative itself
    tau_tot = O.compute_resultant_torques(om, od.joint_positions, od.joint_velocities, tau)
    ref_k = O.system_dynamics(om, od, np.zeros((B, om.number_of_links(), 6)), tau_tot)
    x = dict(base_position=pd._base_position, base_quaternion=pd._base_quaternion, joint_positions=pd._joint_positions,
             base_linear_velocity=pd._base_linear_velocity, base_angular_velocity=pd._base_angular_velocity,
             joint_velocities=pd._joint_velocities, contact_state=pd.contact_state)
    tt = js.ode.compute_resultant_torques(model, pd, joint_force_references=t(tau))
    if om.dofs():
        assert H.rel_err(tt.cpu().numpy(), tau_tot) <= H.RTOL[dtype]
    k = js.ode.system_dynamics(model, x, joint_torques=tt)
    rt = H.RTOL[dtype]
    for key in ("base_position", "base_quaternion", "joint_positions", "base_linear_velocity", "base_angular_velocity", "joint_velocities"):
        if ref_k[key].size:
            assert H.rel_err(k[key].cpu().numpy(), ref_k[key]) <= rt, key
    if ref_k["tangential_deformation"].size:
        assert H.rel_err(k["contact_state"]["tangential_deformation"].cpu().numpy(), ref_k["tangential_deformation"]) <= rt
    # one RK4 step through the public step()
    ref = O.step_rk4(om, od, joint_force_references=tau)
    out = js.model.step(model, pd, joint_force_references=t(tau))
    H.compare_data(out, ref, rt, f"rk4 {name} {dtype}")


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name,B", [("icub_like", 4096), ("ergocub_like", 16384), ("icub_like", 65536)])
def test_full_size_parity(name, B, dtype, cuda_device):
    """BASELINE.json's FULL batch sizes (configs 1, 2, 3): every environment of the batch
    against the plain-C oracle (itself pinned to the NumPy oracle in tests/test_c_oracle.py),
    half of the batch in the free-flight distribution and half touching the ground, plus two
    size-independent properties: unit quaternions and RNEA(ABA(tau)) == tau."""
    import torch

    from oracle import c_oracle as CO

    model = H.build_model(name)
    om = H.oracle_model(model)
    a = O.random_model_data(om, B // 2, seed=61)
    b = O.random_model_data(om, B - B // 2, seed=62, in_contact=True)
    od = O.OracleData(**{f.name: np.concatenate([getattr(a, f.name), getattr(b, f.name)], axis=0)
                         for f in __import__("dataclasses").fields(O.OracleData)})
    rng = np.random.default_rng(11)
    tau = 10 * rng.uniform(size=(B, om.dofs()))
    ref = CO.step(om, od, joint_force_references=tau)
    td = _dtype(dtype)
    pd = H.to_product(model, od, td, cuda_device)
    tau_t = torch.as_tensor(tau, dtype=td, device=cuda_device)
    out = js.model.step(model, pd, joint_force_references=tau_t)
    H.compare_data(out, ref, H.RTOL[dtype], f"full size {name} B={B} {dtype}")
    qn = torch.linalg.norm(out.base_quaternion, dim=-1)
    assert float((qn - 1).abs().max()) <= (1e-12 if dtype == "float64" else 1e-6)
    acc, sdd = js.model.forward_dynamics_aba(model, pd, joint_forces=tau_t)
    _, tau_id = js.model.inverse_dynamics(model, pd, joint_accelerations=sdd, base_acceleration=acc)
    assert float((tau_id - tau_t).abs().max()) / 10.0 <= (1e-9 if dtype == "float64" else 2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("contact", ["soft", "rigid"])
def test_rollout_records_the_trajectory(contact, cuda_device):
    """js.model.rollout (trajectory subsampling): sample t is the state after (t+1) * record_every steps,
    identical to stepping one at a time; the final data carries its caches."""
    import torch

    from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams

    kw = dict(contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build()) if contact == "rigid" else {}
    model = H.build_model("icub_like", **kw)
    om = H.oracle_model(model)
    B, T, k = 9, 12, 3
    od = O.random_model_data(om, B, seed=31, in_contact="flat" if contact == "rigid" else True)
    data = H.to_product(model, od, torch.float64, cuda_device)
    tau = torch.as_tensor(np.random.default_rng(2).uniform(-3, 3, (T, B, om.dofs())), device=cuda_device)
    final, traj = js.model.rollout(model, data, T, joint_force_references=tau, record_every=k)
    assert traj.joint_positions.shape == (T // k, B, om.dofs()) and traj._link_transforms is None
    cur, samples = data, []
    for t in range(T):
        cur = js.model.step(model, cur, joint_force_references=tau[t])
        if (t + 1) % k == 0:
            samples.append(cur)
    for i, ref in enumerate(samples):
        for _, leaf in H.LEAVES[:6]:
            got, want = getattr(traj, leaf)[i], getattr(ref, leaf)
            # (a fused launch carries the kinematics on chip, single steps re-read them from the caches: rounding-level)
            assert torch.allclose(got, want, rtol=1e-9, atol=1e-11), (i, leaf)
    for _, leaf in H.LEAVES:
        assert torch.allclose(getattr(final, leaf), getattr(samples[-1], leaf), rtol=1e-9, atol=1e-11), leaf
    if contact == "soft":
        assert torch.allclose(traj.contact_state["tangential_deformation"][-1], samples[-1].contact_state["tangential_deformation"],
                              rtol=1e-9, atol=1e-13)
    with pytest.raises(ValueError):
        js.model.rollout(model, data, 10, record_every=3)
