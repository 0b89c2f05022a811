"""CPU tests that PIN the oracle (oracle/jaxsim_oracle.py).

The reference's tests hold no golden vectors for this path; they pin it by
known answers and invariants.  Each test below restates one of them against the oracle:

* ``tests/test_api_model.py:495-577``  ABA == CRB forward dynamics, RNEA(ABA(tau)) == tau;
* ``tests/test_simulations.py:15-85``   box held by a balancing wrench does not move;
* ``tests/test_simulations.py:88-167``  zero-gravity box: p = p0 + 1/2 f/m t^2 (atol 1e-3);
* ``tests/test_simulations.py:194-242`` soft-contact rest height z + delta_max == h/2;
* ``tests/test_simulations.py:347-401`` joint-limit spring keeps the joint at the limit;
* ``tests/test_actuation.py:11-48``     torque-speed curve;
plus closed-form answers (pendulum acceleration, free fall) that do not depend on any code
of ours.  Default tolerance of the reference's ``assert_allclose``: rtol 1e-7, atol 1e-9
(``tests/utils.py:14-26``).
"""

import numpy as np
import pytest

from jaxsim_b200 import models
from jaxsim_b200.parsers.urdf import build_kin_dyn_parameters
from oracle import jaxsim_oracle as O

ALL = ["pendulum", "double_pendulum", "cartpole", "box", "sphere", "icub_like", "ergocub_like"]


def omodel(name, **kw):
    _, kd, fb = build_kin_dyn_parameters(models.urdf(name))
    return O.OracleModel(kin_dyn_parameters=kd, floating_base=fb, **kw)


def state(d):
    return (d.base_position, d.base_orientation, d.joint_positions, d.base_linear_velocity,
            d.base_angular_velocity, d.joint_velocities)


@pytest.mark.parametrize("name", ALL)
def test_fd_id_consistency(name):
    m = omodel(name)
    B = 6
    d = O.random_model_data(m, B, seed=1)
    rng = np.random.default_rng(0)
    tau = 10 * rng.uniform(size=(B, m.dofs()))
    W_f = rng.uniform(size=(B, m.number_of_links(), 6))
    if not m.floating_base:
        W_f[:, 0] = 0
    a_aba, sdd_aba = O.aba(m, *state(d), tau, W_f)
    a_crb, sdd_crb = O.forward_dynamics_crb(m, *state(d), tau, W_f)
    np.testing.assert_allclose(sdd_aba, sdd_crb, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(a_aba, a_crb, rtol=1e-7, atol=1e-9)
    fB, tau_id = O.rnea(m, *state(d), a_aba, sdd_aba, W_f)
    np.testing.assert_allclose(tau_id, tau, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(fB, 0.0, atol=1e-8)
    if m.floating_base:
        W_f2 = W_f.copy()
        W_f2[:, 0] = 0
        fB, tau_id = O.rnea(m, *state(d), a_aba, sdd_aba, W_f2)
        np.testing.assert_allclose(tau_id, tau, rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(fB, W_f[:, 0], rtol=1e-7, atol=1e-8)


def test_box_with_external_forces():
    m = omodel("box")
    d0 = O.data_replace(m, np.zeros((1, 0)), np.zeros((1, 0)), np.array([[1.0, 0, 0, 0]]), np.zeros((1, 3)),
                        np.zeros((1, 3)), np.array([[0.0, 0.0, 0.5]]))
    mg = -m.gravity * m.kin_dyn_parameters.link_parameters.mass.sum()
    L_f = np.array([[[0.0, 0.0, mg, 0, 0, 0]]])  # CoM == link origin for this box
    d = d0
    for _ in range(500):
        W_f = O.other_representation_to_inertial(L_f, "body", d.link_transforms, is_force=True)
        d = O.step(m, d, link_forces_inertial=W_f)
    np.testing.assert_allclose(d.base_position, d0.base_position, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(d.base_orientation, d0.base_orientation, rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("repr_", ["inertial", "body", "mixed"])
def test_box_with_zero_gravity(repr_):
    m = omodel("box", gravity=0.0, terrain_height=-1e9)
    rng = np.random.default_rng(3)
    p0 = rng.uniform(size=(1, 3))
    d0 = O.data_replace(m, np.zeros((1, 0)), np.zeros((1, 0)), np.array([[1.0, 0, 0, 0]]), np.zeros((1, 3)),
                        np.zeros((1, 3)), p0)
    LW_f = np.zeros((1, 1, 6))
    LW_f[..., 0:3] = 10.0 * rng.uniform(size=3)
    tf, d = 0.01, d0
    for _ in range(10):
        # the force is defined in Mixed representation; hand it over in `repr_`
        W_f = O.other_representation_to_inertial(LW_f, "mixed", d.link_transforms, is_force=True)
        O_f = O.inertial_to_other_representation(W_f, repr_, d.link_transforms, is_force=True)
        W_f2 = O.other_representation_to_inertial(O_f, repr_, d.link_transforms, is_force=True)
        d = O.step(m, d, link_forces_inertial=W_f2)
    mass = m.kin_dyn_parameters.link_parameters.mass.sum()
    np.testing.assert_allclose(d.base_position, p0 + 0.5 * LW_f[:, 0, 0:3] / mass * tf**2, atol=1e-3)


def test_simulation_with_soft_contacts_rest_height():
    # estimate_good_contact_parameters(nc=4, mu=1, xi=1, delta_max=1e-3)
    # (api/contact.py:160-211 -> rbda/contacts/common.py:88-168)
    m = omodel("box")
    mass = m.kin_dyn_parameters.link_parameters.mass.sum()
    dmax, p = 1e-3, 0.5
    K = min(mass * 9.81 / 4 / dmax ** (1 + p), 1e6)
    D = min(1.0 * 2 * np.sqrt(K * mass), 1e4)
    m.K, m.D, m.mu = K, D, 1.0
    enabled = np.zeros(8, dtype=bool)
    enabled[[0, 1, 2, 3]] = True
    m.kin_dyn_parameters.contact_parameters.enabled = tuple(enabled.tolist())
    box_height = 0.1
    d = O.data_replace(m, np.zeros((1, 0)), np.zeros((1, 0)), np.array([[1.0, 0, 0, 0]]), np.zeros((1, 3)),
                       np.zeros((1, 3)), np.array([[0.0, 0.0, 2 * box_height]]))
    d0 = d
    for _ in range(1000):
        d = O.step(m, d)
    np.testing.assert_allclose(d.base_position[:, 0:2], d0.base_position[:, 0:2], atol=1e-9)
    np.testing.assert_allclose(d.base_position[:, 2] + dmax, box_height / 2, rtol=1e-6, atol=1e-8)


def test_joint_limits_spring():
    # gravity pointing UP turns the hanging pendulum into the reference's inverted one
    # (tests/conftest.py:370-476): it falls away from s = 0 into the joint limit.
    m = omodel("pendulum", time_step=0.001, gravity=+9.81)
    jp = m.kin_dyn_parameters.joint_parameters
    jp.position_limits_max = np.array([1.5708])
    jp.position_limits_min = np.array([-1.5708])
    jp.position_limit_spring = np.array([75.0])
    jp.position_limit_damper = np.array([0.1])
    jp.friction_viscous = np.array([0.5])  # let it settle
    theta = 10 * np.pi / 180
    z3 = np.zeros((1, 3))
    d = O.data_replace(m, jp.position_limits_min[None] - theta, np.zeros((1, 1)), np.array([[1.0, 0, 0, 0]]), z3, z3, z3)
    smin_seen = np.inf
    for _ in range(6000):
        d = O.step(m, d)
        smin_seen = min(smin_seen, d.joint_positions.min())
    # steady state: the limit torque balances gravity (joint beyond the limit by tau_g / k)
    s = d.joint_positions[0, 0]
    assert abs(d.joint_velocities[0, 0]) < 1e-3
    tau_lim = -75.0 * min(s - (-1.5708), 0.0)
    _, tau_g = O.rnea(m, d.base_position, d.base_orientation, d.joint_positions, z3, z3, np.zeros((1, 1)),
                      np.zeros((1, 6)), np.zeros((1, 1)), np.zeros((1, 2, 6)))
    np.testing.assert_allclose(tau_lim, tau_g[0, 0], rtol=1e-3)
    assert s < -1.5708 and s > -1.5708 - theta  # pushed back towards the limit


def test_actuation_damper_expression():
    """``tau_limit -= (positive(tau_limit) * diag(d)) @ sd`` == tau_limit (1 - d sd)
    (api/actuation_model.py:64-66) -- hand-computed numbers."""
    m = omodel("pendulum")
    jp = m.kin_dyn_parameters.joint_parameters
    jp.position_limits_max = np.array([1.0])
    jp.position_limits_min = np.array([-1.0])
    jp.position_limit_spring = np.array([100.0])
    jp.position_limit_damper = np.array([0.5])
    m.enable_friction = False
    s = np.array([[1.2]])
    sd = np.array([[0.4]])
    tau = O.compute_resultant_torques(m, s, sd, np.zeros((1, 1)))
    # spring: -100 * 0.2 = -20 ; damper quirk: -20 - (-20 * 0.5 * 0.4) = -16
    np.testing.assert_allclose(tau, [[-16.0]], rtol=1e-12)


def test_tn_curve():
    m = omodel("pendulum", torque_max=10.0, omega_th=1.0, omega_max=2.0)
    s = np.zeros((1, 1))
    tau0 = 30 * np.ones((1, 1))
    tau = O.compute_resultant_torques(m, s, 1.5 * np.ones((1, 1)), tau0)
    assert np.all(tau < tau0)
    np.testing.assert_allclose(tau, 5.0)  # 10 * (1 - 0.5 / 1)
    tau = O.compute_resultant_torques(m, s, 2.5 * np.ones((1, 1)), tau0)
    np.testing.assert_allclose(tau, 0.0)


def test_pendulum_closed_form():
    """sdd = -m g l sin(s) / I_pivot for a pendulum about x hanging along -z."""
    m = omodel("pendulum")
    kd = m.kin_dyn_parameters
    M = kd.link_parameters.spatial_inertias()[1]
    mass = kd.link_parameters.mass[1]
    com = kd.link_parameters.center_of_mass[1]
    I_xx = M[3, 3]  # about the link origin == joint axis
    z3 = np.zeros((1, 3))
    for s in (0.3, -1.1, 2.0):
        _, sdd = O.aba(m, z3, np.array([[1.0, 0, 0, 0]]), np.array([[s]]), z3, z3, np.zeros((1, 1)),
                       np.zeros((1, 1)), np.zeros((1, 2, 6)))
        # gravity torque about x of a mass at R_x(s) com
        c, sn = np.cos(s), np.sin(s)
        y = c * com[1] - sn * com[2]
        expected = (y * mass * m.gravity) / I_xx
        np.testing.assert_allclose(sdd[0, 0], expected, rtol=1e-10)


def test_free_fall_closed_form():
    m = omodel("box", terrain_height=-1e9)
    d = O.random_model_data(m, 5, seed=2)
    W_f = np.zeros((5, 1, 6))
    a, _ = O.aba(m, *state(d), np.zeros((5, 0)), W_f)
    # inertial-fixed base acceleration of a free body: lin = g + (v_lin x w ... ) terms.
    # For a body spinning freely, the CoM (== link origin here) accelerates with g only:
    # d/dt(p) = v_lin + w x p  =>  pdd = a_lin + alpha x p + w x pd
    pd = d.base_linear_velocity + np.cross(d.base_angular_velocity, d.base_position)
    pdd = a[:, 0:3] + np.cross(a[:, 3:6], d.base_position) + np.cross(d.base_angular_velocity, pd)
    np.testing.assert_allclose(pdd, np.tile([0, 0, m.gravity], (5, 1)), rtol=1e-9, atol=1e-9)


def test_step_quaternion_stays_unit_and_caches_consistent():
    m = omodel("icub_like")
    d = O.random_model_data(m, 4, seed=5, in_contact=True)
    for _ in range(5):
        d = O.step(m, d)
    np.testing.assert_allclose(np.linalg.norm(d.base_quaternion, axis=-1), 1.0, atol=1e-14)
    d2 = O.data_replace(m, d.joint_positions, d.joint_velocities, d.base_quaternion, d.base_linear_velocity,
                        d.base_angular_velocity, d.base_position)
    np.testing.assert_allclose(d.link_transforms, d2.link_transforms, atol=1e-14)
    np.testing.assert_allclose(d.link_transforms[:, 0], d.base_transform, atol=1e-14)


def test_in_contact_distribution_touches_ground():
    m = omodel("icub_like")
    d = O.random_model_data(m, 32, seed=9, in_contact=True)
    W_p_C, _ = O.collidable_points_pos_vel(m, d.link_transforms, d.link_velocities)
    zmin = W_p_C[..., 2].min(axis=1)
    assert np.all(zmin <= 0.0) and np.all(zmin >= -0.0051)
    W_f, _ = O.soft_compute_contact_forces(m, d)
    assert np.any(np.abs(W_f) > 0)
