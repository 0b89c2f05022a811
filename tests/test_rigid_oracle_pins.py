"""Pins of ``oracle/rigid_oracle.py`` (CPU): the reference's known-answer test for the
rigid-contact path plus the invariants that make its pieces unambiguous."""

import dataclasses

import numpy as np
import pytest

from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams
from oracle import jaxsim_oracle as O
from oracle import rigid_oracle as R

from . import helpers as H


def _model(name, **params):
    return H.build_model(name, contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build(**params))


def test_box_rests_at_half_height():
    """``tests/test_simulations.py:245-292``: K=1e5, the 4 bottom corners enabled, dropped from
    2h, after 1 s: x,y unchanged and z == h/2 (assert_allclose default rtol 1e-7)."""
    model = _model("box", K=1e5)
    cp = model.kin_dyn_parameters.contact_parameters
    model.kin_dyn_parameters.contact_parameters = dataclasses.replace(cp, enabled=tuple([True] * 4 + [False] * 4))
    om = H.oracle_model(model)
    d = O.data_replace(om, np.zeros((1, 0)), np.zeros((1, 0)), np.array([[1.0, 0, 0, 0]]), np.zeros((1, 3)),
                       np.zeros((1, 3)), np.array([[0.0, 0.0, 0.2]]))
    for _ in range(1000):
        d = R.step(om, d)
    np.testing.assert_allclose(d.base_position[0, 0:2], 0.0, atol=1e-12)
    np.testing.assert_allclose(d.base_position[0, 2], 0.05, rtol=1e-7)


@pytest.mark.parametrize("name", ["box", "icub_like"])
def test_mass_inverse_and_jacobians(name):
    """M^-1 (rbda/mass_inverse.py) inverts the CRBA mass matrix; J nu is the point velocity;
    J_dot nu + J nu_dot is the classical point acceleration omega x pdot + a + alpha x rho."""
    model = _model(name)
    om = H.oracle_model(model)
    d = O.random_model_data(om, 2, seed=4, in_contact=True)
    env = R._Env(d, 1)
    M, Mi = R.mass_matrix_mixed(om, env), R.mass_matrix_inverse_mixed(om, env)
    np.testing.assert_allclose(Mi @ M, np.eye(M.shape[0]), atol=1e-9)
    n = om.dofs()
    J = R.contact_jacobian_mixed(om, env, "mixed")
    nu = env.generalized_velocity("mixed")
    _, W_pd = O.collidable_points_pos_vel(om, env.W_H_L[None], env.W_v_WL[None])
    np.testing.assert_allclose((J.reshape(-1, 6 + n) @ nu).reshape(-1, 6)[:, 0:3], W_pd[0], atol=1e-12)
    # finite-difference check of J_dot: d/dt (J nu) along the free motion
    Jd = R.contact_jacobian_derivative_mixed(om, env)
    eps = 1e-6
    s2 = d.joint_positions[1] + eps * d.joint_velocities[1]
    W_pd_B = env.base_velocity("mixed")[0:3]
    p2 = d.base_position[1] + eps * W_pd_B
    w = d.base_angular_velocity[1]
    q = d.base_quaternion[1]
    dq = 0.5 * O._qmul(np.array([[0.0, *w]]), q[None])[0]
    q2 = q + eps * dq
    # same mixed generalized velocity at the displaced configuration
    v_lin2 = W_pd_B - np.cross(w, p2)
    d2 = O.data_replace(om, s2[None], d.joint_velocities[1][None], q2[None], v_lin2[None], w[None], p2[None])
    env2 = R._Env(d2, 0)
    J2 = R.contact_jacobian_mixed(om, env2, "mixed")
    fd = ((J2 - J).reshape(-1, 6 + n) @ nu / eps).reshape(-1, 6)[:, 0:3]
    an = (Jd.reshape(-1, 6 + n) @ nu).reshape(-1, 6)[:, 0:3]
    np.testing.assert_allclose(an, fd, atol=2e-5 * max(1.0, np.abs(fd).max()))


def test_qp_solver_reaches_the_kkt_point():
    """A KKT point of a strictly convex QP is its unique optimum: check stationarity, primal
    and dual feasibility and complementarity of ``solve_qp`` on contact-like problems, and
    agreement with scipy's SLSQP."""
    from scipy.optimize import minimize

    rng = np.random.default_rng(0)
    for trial in range(6):
        na = 1 + trial % 4
        A = rng.normal(size=(3 * na, 4))
        Q = A @ A.T + 1e-6 * np.eye(3 * na)  # rank deficient + regularisation, like the Delassus matrix
        q = A @ rng.normal(size=4) * 10 + 1e-5 * rng.normal(size=3 * na)  # (almost) in range(Q), like J(...)
        G = R.ineq_constraint_matrix(np.zeros(na, dtype=bool), 0.5)
        h = np.zeros(G.shape[0])
        x, s, z, conv, it = R.solve_qp(Q, q, G, h)
        assert conv
        assert np.abs(Q @ x + q + G.T @ z).max() < 1e-8 * (1 + np.abs(q).max() + np.abs(Q @ x).max())
        assert (G @ x <= 1e-9).all() and (z >= -1e-12).all()
        assert np.abs(z * (G @ x)).max() < 1e-8 * (1 + np.abs(q).max())
        res = minimize(lambda y: 0.5 * y @ Q @ y + q @ y, np.zeros(3 * na), jac=lambda y: Q @ y + q, method="SLSQP",
                       constraints=[{"type": "ineq", "fun": lambda y: -(G @ y), "jac": lambda y: -G}],
                       options={"ftol": 1e-14, "maxiter": 500})
        f = lambda y: 0.5 * y @ Q @ y + q @ y  # noqa: E731
        assert f(x) <= f(res.x) + 1e-7 * (1 + abs(f(res.x)))


def test_inactive_points_carry_no_force_and_impact_stops_active_points():
    model = _model("icub_like", K=1e3, D=10.0)
    om = H.oracle_model(model)
    d = O.random_model_data(om, 2, seed=9, in_contact="flat")
    n, nL = om.dofs(), om.number_of_links()
    env = R._Env(d, 0)
    W_f_C = R.compute_contact_forces(om, env, np.zeros(n), np.zeros((nL, 6)))
    W_p_C, _ = O.collidable_points_pos_vel(om, d.link_transforms, d.link_velocities)
    inactive = W_p_C[0][:, 2] >= 0
    assert inactive.any() and (~inactive).sum() >= 2
    assert np.abs(W_f_C[inactive]).max() == 0.0
    assert (W_f_C[~inactive][:, 2] >= -1e-9).all()
    d2 = R.step(om, d)
    env2 = R._Env(d2, 0)
    J2 = R.contact_jacobian_mixed(om, env2, "mixed")
    nu2 = env2.generalized_velocity("mixed")
    W_p_C2, _ = O.collidable_points_pos_vel(om, d2.link_transforms, d2.link_velocities)
    act2 = W_p_C2[0][:, 2] < 0
    v = (J2.reshape(-1, 6 + n) @ nu2).reshape(-1, 6)[:, 0:3]
    assert np.abs(v[act2]).max() < 1e-10


def test_rigid_contacts_host_objects():
    cm = RigidContacts.build(regularization_delassus=1e-5, solver_options={"solver_tol": 1e-4})
    assert cm.regularization_delassus == 1e-5 and dict(cm.solver_options)["solver_tol"] == 1e-4
    with pytest.raises(ValueError):
        RigidContacts.build(solver_options={"x": []})
    prm = RigidContactsParams.build(mu=0.7)
    assert prm.valid() and prm.K == 0.0 and prm.D == 0.0 and prm.mu == 0.7
    model = H.build_model("box", contact_model=RigidContacts.build())
    assert isinstance(model.contact_params, RigidContactsParams)
