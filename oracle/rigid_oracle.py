"""CPU oracle for the RigidContacts variant of ``jaxsim.api.model.step`` (BASELINE config 3).

TEST INFRASTRUCTURE ONLY (same rules as ``jaxsim_oracle.py``: imported by ``tests/``,
``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` only).

Literal NumPy restatement, one environment at a time (Python loop over the batch), of

* ``RigidContacts.compute_contact_forces`` / ``update_velocity_after_impact`` /
  ``compute_impact_velocity`` and the helpers ``_delassus_matrix``,
  ``_compute_ineq_constraint_matrix``, ``_linear_acceleration_of_collidable_points``,
  ``_compute_baumgarte_stabilization_term``           (``rbda/contacts/rigid.py:163-539``)
* ``js.contact.transforms / jacobian / jacobian_derivative``  (``api/contact.py:214-511``)
* ``generalized_free_floating_jacobian(_derivative)``          (``api/model.py:925-1228``)
* ``jacobian_full_doubly_left`` / ``jacobian_derivative_full_doubly_left``
                                                                (``rbda/jacobian.py:128-339``)
* ``mass_inverse``                                              (``rbda/mass_inverse.py:11-233``)
* ``free_floating_mass_matrix(_inverse)`` + ``_transform_M_block`` (``api/model.py:1529-1631``)
* ``forward_dynamics_aba``'s ``to_active`` conversion           (``api/model.py:1356-1404``)
* the wiring in ``ode.system_acceleration`` (``api/ode.py:16-131``) and ``step``
  (``api/model.py:2601-2681``).

Third-party arithmetic: the QP ``min 1/2 x'Qx + q'x  s.t. Gx <= h`` is solved in the reference by
``qpax.solve_qp`` (``pyproject.toml:55``, NO version pin, not vendored) with ``solver_tol=1e-3``
(``rigid.py:99-108,360-362``).  qpax is a primal-dual interior-point method (Mehrotra
predictor-corrector, after Mattingley & Boyd's CVXGEN).  Its iterates cannot be reproduced
without its source, and at ``solver_tol=1e-3`` they are only an approximation of the optimum,
so -- as SURVEY.md 8c prescribes -- parity is defined on THE optimum of the QP, which is unique
because ``Q = J M^-1 J' + 1e-6 I`` is positive definite.  ``solve_qp`` below is an independent
restatement of the published algorithm run to tight tolerance; ``tests/test_oracle_pins.py``
pins it by checking the KKT conditions of its answers (a KKT point of a strictly convex QP is
the optimum) and against ``scipy.optimize`` on small problems.

PARITY PIN STATUS: pinned by the reference's own code up to the QP boundary, and on the QP
optimum beyond it.  The unmodified reference sources (``rbda/contacts/rigid.py``,
``api/contact.py``, ``rbda/mass_inverse.py``, ``rbda/jacobian.py`` ...) are executed over NumPy
stand-ins (``oracle/refshim``; its ``qpax.solve_qp`` is a separate interior-point code run to
tight tolerance) by ``tests/golden/make_goldens.py``; ``tests/test_reference_goldens.py`` checks
every leaf of this oracle's ``step`` against those fixtures at 1e-6 relative.  The path is
also pinned by the reference's known-answer test for it: a box dropped on the ground comes to
rest at height h/2 with ~zero penetration (``tests/test_simulations.py:245-292``), re-run
against this restatement, plus the invariants J M^-1 J' == Delassus from ABA impulse
responses, M^-1 M == I, and J(post-impact nu) == 0 on the active points.

Quirks reproduced on purpose:
* ``update_velocity_after_impact`` replaces the state velocities with ``dataclasses.replace``
  (``rigid.py:429-434``), i.e. WITHOUT refreshing the cached ``_link_velocities``: the caches of
  the returned data hold the PRE-impact link velocities, and the next step's penetration rate
  and Jacobian-derivative term ``O_X_dot_W`` (``api/contact.py:470-477``) read those.
* ``jacobian(model, data)[indices_of_enabled_collidable_points]`` and the same for the point
  positions (``rigid.py:401-409``) index the already-filtered arrays a second time (restated
  with JAX's clamping of out-of-range gather indices).  It is the identity when the enabled
  points are a prefix of the point list -- the only case the reference's own test uses
  (``tests/test_simulations.py:262-268``) and the only one the CUDA path accepts.
"""

from __future__ import annotations

import numpy as np

from . import jaxsim_oracle as O

# ---------------------------------------------------------------------------------------
# rbda/jacobian.py:128-339
# ---------------------------------------------------------------------------------------


def _joint_transforms_1(model, s, W_H_B):
    return O.joint_transforms(model, s[None], W_H_B[None])[0]


def jacobian_full_doubly_left(model, s):
    """``rbda/jacobian.py:128-212`` -> (B_J_full (6,6+n), B_H_L (nL,4,4))."""
    kd = model.kin_dyn_parameters
    lam = kd.parent_array
    nL, n = kd.number_of_links(), kd.number_of_joints()
    dt = s.dtype
    i_X_lam = _joint_transforms_1(model, s, np.eye(4, dtype=dt))
    S = kd.motion_subspaces.astype(dt)
    B_X_i = np.zeros((nL, 6, 6), dtype=dt)
    B_X_i[0] = np.eye(6)
    J = np.zeros((6, 6 + n), dtype=dt)
    J[0:6, 0:6] = np.eye(6)
    for i in range(1, nL):
        B_X_i[i] = B_X_i[lam[i]] @ O.adjoint_inverse(i_X_lam[i])
        J[:, 6 + i - 1] = B_X_i[i] @ S[i]
    return J, O.adjoint_to_transform(B_X_i)


def jacobian_derivative_full_doubly_left(model, s, sd):
    """``rbda/jacobian.py:215-339`` -> (B_Jdot_full (6,6+n), B_H_L)."""
    kd = model.kin_dyn_parameters
    lam = kd.parent_array
    nL, n = kd.number_of_links(), kd.number_of_joints()
    dt = s.dtype
    i_X_lam = _joint_transforms_1(model, s, np.eye(4, dtype=dt))
    S = kd.motion_subspaces.astype(dt)
    B_X_i = np.zeros((nL, 6, 6), dtype=dt)
    B_X_i[0] = np.eye(6)
    B_Xd_i = np.zeros((nL, 6, 6), dtype=dt)
    B_v_Bi = np.zeros((nL, 6), dtype=dt)
    Jd = np.zeros((6, 6 + n), dtype=dt)
    for i in range(1, nL):
        ii = i - 1
        B_X_i[i] = B_X_i[lam[i]] @ O.adjoint_inverse(i_X_lam[i])
        B_v_Bi[i] = B_v_Bi[lam[i]] + B_X_i[i] @ S[i] * sd[ii]
        i_X_B = O.adjoint_inverse(B_X_i[i])
        B_Xd_i[i] = B_X_i[i] @ O.cross_vx(i_X_B @ B_v_Bi[i])
        Jd[:, 6 + ii] = B_Xd_i[i] @ S[i]
    return Jd, O.adjoint_to_transform(B_X_i)


# ---------------------------------------------------------------------------------------
# per-environment view of the data + representation helpers (api/data.py:288-341)
# ---------------------------------------------------------------------------------------


class _Env:
    """One environment of an ``OracleData`` (state + caches)."""

    def __init__(self, data, b):
        self.s = data.joint_positions[b]
        self.sd = data.joint_velocities[b]
        self.q = data.base_orientation[b]
        self.p = data.base_position[b]
        self.W_v_WB = np.concatenate([data.base_linear_velocity[b], data.base_angular_velocity[b]])
        self.W_H_B = data.base_transform[b]
        self.W_H_L = data.link_transforms[b]
        self.W_v_WL = data.link_velocities[b]

    def base_velocity(self, repr_):
        """``JaxSimModelData.base_velocity`` (``api/data.py:288-312``)."""
        return O.inertial_to_other_representation(self.W_v_WB, repr_, self.W_H_B, is_force=False)

    def generalized_velocity(self, repr_):
        return np.concatenate([self.base_velocity(repr_), self.sd])


def _blockdiag(X, n):
    T = np.zeros((6 + n, 6 + n), dtype=X.dtype)
    T[0:6, 0:6] = X
    T[6:, 6:] = np.eye(n, dtype=X.dtype)
    return T


def _kappa_mask(model, dt):
    kb = model.kin_dyn_parameters.support_body_array_bool
    return np.concatenate([np.ones((kb.shape[0], 5), dtype=dt), kb.astype(dt)], axis=1)  # (nL, 6+n) -- sic: 5 + (n+1)


def generalized_free_floating_jacobian(model, env: _Env, data_repr: str, output_vel_repr: str):
    """``api/model.py:925-1043`` for ``output_vel_repr == "inertial"`` (the only one the
    contact Jacobians request, ``api/contact.py:286-288,447-459``)."""
    assert output_vel_repr == "inertial"
    n = model.dofs()
    dt = env.s.dtype
    B_J_full, _ = jacobian_full_doubly_left(model, env.s)
    if data_repr == "inertial":
        B_X_W = O.adjoint_from_transform(env.W_H_B, inverse=True)
        B_J_full_I = B_J_full @ _blockdiag(B_X_W, n)
    elif data_repr == "body":
        B_J_full_I = B_J_full
    else:  # mixed
        BW_H_B = np.eye(4, dtype=dt)
        BW_H_B[0:3, 0:3] = O.quat_to_dcm(env.q)
        B_X_BW = O.adjoint_from_transform(BW_H_B, inverse=True)
        B_J_full_I = B_J_full @ _blockdiag(B_X_BW, n)
    mask = _kappa_mask(model, dt)  # (nL, 6+n)
    B_J_WL_I = mask[:, None, :] * B_J_full_I[None]
    W_X_B = O.adjoint_from_transform(env.W_H_B)
    return np.einsum("ij,ljk->lik", W_X_B, B_J_WL_I)


def generalized_free_floating_jacobian_derivative_inertial(model, env: _Env):
    """``api/model.py:1046-1228`` with data and output representation both Inertial (what
    ``api/contact.py:447-459`` asks for)."""
    n = model.dofs()
    dt = env.s.dtype
    B_Jd_full, _ = jacobian_derivative_full_doubly_left(model, env.s, env.sd)
    B_J_full, _ = jacobian_full_doubly_left(model, env.s)
    mask = _kappa_mask(model, dt)
    B_Jd_WL_B = mask[:, None, :] * B_Jd_full[None]
    B_J_WL_B = mask[:, None, :] * B_J_full[None]
    W_H_B = env.W_H_B
    # input representation: Inertial
    B_X_W = O.adjoint_from_transform(W_H_B, inverse=True)
    W_v_WB = env.base_velocity("inertial")
    B_Xd_W = -B_X_W @ O.cross_vx(W_v_WB)
    T = _blockdiag(B_X_W, n)
    Td = np.zeros_like(T)
    Td[0:6, 0:6] = B_Xd_W
    # output representation: Inertial
    W_X_B = O.adjoint_from_transform(W_H_B)
    B_v_WB = env.base_velocity("body")
    W_Xd_B = W_X_B @ O.cross_vx(B_v_WB)
    out = np.einsum("ij,ljk->lik", W_Xd_B, B_J_WL_B) @ T
    out = out + np.einsum("ij,ljk->lik", W_X_B, B_Jd_WL_B) @ T
    out = out + np.einsum("ij,ljk->lik", W_X_B, B_J_WL_B) @ Td
    return out


# ---------------------------------------------------------------------------------------
# api/contact.py:214-511
# ---------------------------------------------------------------------------------------


def _enabled(model):
    cp = model.kin_dyn_parameters.contact_parameters
    idx = np.asarray(cp.indices_of_enabled_collidable_points)
    body = np.array(cp.body, dtype=int)[idx]
    return idx, body, np.asarray(cp.point)[idx]


def contact_transforms(model, env: _Env):
    """``api/contact.py:214-256``."""
    _, body, L_p = _enabled(model)
    dt = env.s.dtype
    L_H_C = np.tile(np.eye(4, dtype=dt), (len(body), 1, 1))
    L_H_C[:, 0:3, 3] = L_p
    return env.W_H_L[body] @ L_H_C


def contact_jacobian_mixed(model, env: _Env, data_repr: str):
    """``api/contact.py:259-345`` with ``output_vel_repr = Mixed``."""
    _, body, _ = _enabled(model)
    W_J_WL = generalized_free_floating_jacobian(model, env, data_repr, "inertial")
    W_J_WC = W_J_WL[body]
    W_H_C = contact_transforms(model, env)
    out = np.zeros_like(W_J_WC)
    for c in range(len(body)):
        W_H_CW = W_H_C[c].copy()
        W_H_CW[0:3, 0:3] = np.eye(3)
        CW_X_W = O.adjoint_from_transform(W_H_CW, inverse=True)
        out[c] = CW_X_W @ W_J_WC[c]
    return out


def contact_jacobian_derivative_mixed(model, env: _Env):
    """``api/contact.py:348-511`` with data representation Mixed and output Mixed."""
    _, body, L_p = _enabled(model)
    n = model.dofs()
    dt = env.s.dtype
    W_H_Li = env.W_H_L
    W_v_WLi = env.W_v_WL  # CACHED link velocities (stale after an impact; see module docstring)
    # input representation: Mixed
    W_H_BW = env.W_H_B.copy()
    W_H_BW[0:3, 0:3] = np.eye(3)
    W_X_BW = O.adjoint_from_transform(W_H_BW)
    BW_v_WB = env.base_velocity("mixed")
    BW_v_W_BW = BW_v_WB.copy()
    BW_v_W_BW[3:6] = 0
    W_Xd_BW = W_X_BW @ O.cross_vx(BW_v_W_BW)
    T = _blockdiag(W_X_BW, n)
    Td = np.zeros_like(T)
    Td[0:6, 0:6] = W_Xd_BW
    # link Jacobians in inertial/inertial
    W_J_WL_W = generalized_free_floating_jacobian(model, env, "inertial", "inertial")
    W_Jd_WL_W = generalized_free_floating_jacobian_derivative_inertial(model, env)
    out = np.zeros((len(body), 6, 6 + n), dtype=dt)
    for c in range(len(body)):
        L_H_C = np.eye(4, dtype=dt)
        L_H_C[0:3, 3] = L_p[c]
        W_H_C = W_H_Li[body[c]] @ L_H_C
        W_H_CW = W_H_C.copy()
        W_H_CW[0:3, 0:3] = np.eye(3)
        CW_X_W = O.adjoint_from_transform(W_H_CW, inverse=True)
        CW_v_WC = CW_X_W @ W_v_WLi[body[c]]
        W_v_W_CW = np.zeros(6, dtype=dt)
        W_v_W_CW[0:3] = CW_v_WC[0:3]
        CW_Xd_W = -CW_X_W @ O.cross_vx(W_v_W_CW)
        out[c] = CW_Xd_W @ W_J_WL_W[body[c]] @ T + CW_X_W @ W_Jd_WL_W[body[c]] @ T + CW_X_W @ W_J_WL_W[body[c]] @ Td
    return out


# ---------------------------------------------------------------------------------------
# rbda/mass_inverse.py:11-233 and api/model.py:1529-1631
# ---------------------------------------------------------------------------------------


def mass_inverse_body(model, env: _Env):
    """``rbda/mass_inverse.py:11-233``: M^-1 in body-fixed representation, ABA-like."""
    kd = model.kin_dyn_parameters
    lam = kd.parent_array
    NB, N = kd.number_of_links(), kd.number_of_joints()
    dt = env.s.dtype
    nv = N + 6
    I_A = O.link_spatial_inertia_matrices(model, dt).copy()
    W_H_B = O.transform_from_quat_pos(env.q[None], env.p[None])[0]
    X = _joint_transforms_1(model, env.s, W_H_B)
    S = kd.motion_subspaces.astype(dt)
    F = np.zeros((NB, 6, nv), dtype=dt)
    P = np.zeros((NB, 6, nv), dtype=dt)
    U = np.zeros((NB, 6), dtype=dt)
    D = np.zeros((NB,), dtype=dt)
    Minv = np.zeros((nv, nv), dtype=dt)
    for i in range(NB - 1, 0, -1):
        Si, Fi, Xi, parent = S[i], F[i].copy(), X[i], lam[i]
        Ui = I_A[i] @ Si
        Di = Si @ Ui
        U[i], D[i] = Ui, Di
        r = 6 + (i - 1)
        row = Minv[r].copy()
        row[r] += 1.0 / Di
        row = row - (Si @ Fi) / Di
        Minv[r] = row
        if parent >= 0:
            Fa_i = Fi + Ui[:, None] @ row[None, :]
            F[parent] = F[parent] + Xi.T @ Fa_i
            Ia_i = I_A[i] - np.outer(Ui, Ui) / Di
            I_A[parent] = I_A[parent] + Xi.T @ Ia_i @ Xi
    D0_inv = np.linalg.inv(I_A[0])
    Minv[0:6, 0:6] += D0_inv
    Minv[0:6, :] += -(D0_inv.T @ F[0])
    P[0] = Minv[0:6, :]
    for i in range(1, NB):
        Si, Ui, Di, Xi, parent = S[i], U[i], D[i], X[i], lam[i]
        P_parent = P[parent] if parent >= 0 else np.zeros_like(P[i])
        r = 6 + (i - 1)
        if parent >= 0:
            Minv[r, :] = Minv[r, :] - (Ui @ (Xi @ P_parent)) / Di
        P[i] = Si[:, None] @ Minv[r, :][None, :] + Xi @ P_parent
    return 0.5 * (Minv + Minv.T)


def _transform_M_block(M_body, X):
    """``api/model.py:1529-1556``."""
    M = M_body.copy()
    M[:6, :6] = X.T @ M_body[:6, :6] @ X
    M[:6, 6:] = X.T @ M_body[:6, 6:]
    M[6:, :6] = M_body[6:, :6] @ X
    return M


def mass_matrix_mixed(model, env: _Env):
    """``free_floating_mass_matrix`` (``api/model.py:1559-1592``) in Mixed."""
    M_body = O.crba(model, env.s[None])[0]
    BW_H_B = env.W_H_B.copy()
    BW_H_B[0:3, 3] = 0
    B_X_BW = O.adjoint_from_transform(BW_H_B, inverse=True)
    return _transform_M_block(M_body, B_X_BW)


def mass_matrix_inverse_mixed(model, env: _Env):
    """``free_floating_mass_matrix_inverse`` (``api/model.py:1595-1631``) in Mixed."""
    Minv_body = mass_inverse_body(model, env)
    B_H_BW = env.W_H_B.copy()
    B_H_BW[0:3, 3] = 0
    BW_X_B = O.adjoint_from_transform(B_H_BW)
    return _transform_M_block(Minv_body, BW_X_B.T)


# ---------------------------------------------------------------------------------------
# QP:  min 1/2 x'Qx + q'x  s.t.  Gx <= h     (qpax.solve_qp restated; see module docstring)
# ---------------------------------------------------------------------------------------


def solve_qp(Q, q, G, h, tol=1e-11, max_iter=200):
    """Primal-dual interior point (Mehrotra predictor-corrector) for an inequality-only,
    strictly convex QP.  Returns (x, s, z, converged, iters).  Rows of G that are identically
    zero with h == 0 (``rigid.py:478-489`` row 5 of active points) are dropped: they
    constrain nothing.  The best iterate (smallest scaled KKT residual) is returned; the
    iteration stops when that residual is below ``tol`` or stalls at rounding level while the
    duality gap is already far below ``tol`` (``converged`` then reports residual < 1e-8)."""
    Q = np.asarray(Q, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    G = np.asarray(G, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    keep = ~((np.abs(G).sum(axis=1) == 0) & (h >= 0))
    Gk, hk = G[keep], h[keep]
    nx, m = Q.shape[0], Gk.shape[0]
    x = np.zeros(nx)
    if m == 0:
        x = np.linalg.solve(Q, -q)
        return x, np.zeros(G.shape[0]), np.zeros(G.shape[0]), True, 0
    s = np.ones(m)
    z = np.ones(m)
    best = (np.inf, x.copy(), s.copy(), z.copy())
    it = 0
    for it in range(1, max_iter + 1):
        Qx = Q @ x
        r_d = Qx + q + Gk.T @ z
        r_p = Gk @ x + s - hk
        mu = s @ z / m
        m_d = np.abs(r_d).max() / (1.0 + np.abs(q).max() + np.abs(Qx).max())
        m_p = np.abs(r_p).max() / (1.0 + np.abs(x).max())
        m_g = mu / (1.0 + abs(0.5 * x @ Qx + q @ x))
        merit = max(m_d, m_p, m_g)
        if merit < best[0]:
            best = (merit, x.copy(), s.copy(), z.copy())
        if not (merit > tol) or not (m_g > 1e-3 * tol) or not np.isfinite(merit):
            break
        W = z / s
        H = Q + Gk.T @ (W[:, None] * Gk)
        try:
            L = np.linalg.cholesky(H)
        except np.linalg.LinAlgError:
            break

        def newton(r_c):
            # r_c: target residual of s*z (to be driven to zero): s*dz + z*ds = -r_c
            rhs = -(r_d + Gk.T @ ((z * r_p - r_c) / s))
            dx = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
            ds = -r_p - Gk @ dx
            dz = -(r_c + z * ds) / s
            return dx, ds, dz

        def max_step(v, dv):
            neg = dv < 0
            return float(np.min(-v[neg] / dv[neg])) if neg.any() else np.inf

        dxa, dsa, dza = newton(s * z)
        a_aff = min(1.0, max_step(s, dsa), max_step(z, dza))
        mu_aff = (s + a_aff * dsa) @ (z + a_aff * dza) / m
        sigma = (mu_aff / mu) ** 3
        dx, ds, dz = newton(s * z + dsa * dza - sigma * mu)
        a = min(1.0, 0.99 * min(max_step(s, ds), max_step(z, dz)))
        x = x + a * dx
        s = s + a * ds
        z = z + a * dz
    merit, x, s, z = best
    s_full = np.zeros(G.shape[0])
    z_full = np.zeros(G.shape[0])
    s_full[keep], z_full[keep] = s, z
    return x, s_full, z_full, bool(merit < 1e-8), it


# ---------------------------------------------------------------------------------------
# rbda/contacts/rigid.py
# ---------------------------------------------------------------------------------------


def ineq_constraint_matrix(inactive, mu):
    """``_compute_ineq_constraint_matrix`` (``rigid.py:472-497``)."""
    nc = len(inactive)
    G1 = np.array([[1, 0, -mu], [0, 1, -mu], [-1, 0, -mu], [0, -1, -mu], [0, 0, -1], [0, 0, 0]], dtype=np.float64)
    G = np.zeros((6 * nc, 3 * nc))
    for c in range(nc):
        Gc = G1.copy()
        Gc[5, 2] = float(inactive[c])
        G[6 * c:6 * c + 6, 3 * c:3 * c + 3] = Gc
    return G


def _solve_contact_qp(Q, q, inactive, mu):
    """The optimum of the reference's QP.  Inactive points are pinned to zero force by rows
    4-5 (f_z >= 0 and f_z <= 0) and the pyramid rows; their feasible set has no interior, so
    they are eliminated exactly before the interior-point solve (same optimum)."""
    nc = len(inactive)
    act = np.where(~inactive)[0]
    x = np.zeros(3 * nc)
    if len(act) == 0:
        return x
    sel = (3 * act[:, None] + np.arange(3)[None, :]).reshape(-1)
    G = ineq_constraint_matrix(np.zeros(len(act), dtype=bool), mu)
    xa, _, _, conv, _ = solve_qp(Q[np.ix_(sel, sel)], q[sel], G, np.zeros(G.shape[0]))
    if not conv:
        raise RuntimeError("rigid oracle: QP did not converge")
    x[sel] = xa
    return x


def aba_mixed(model, env: _Env, tau, W_f_L):
    """``forward_dynamics_aba`` with the data in Mixed representation (``api/model.py:1269-1406``):
    inertial ABA, then ``to_active`` for the base acceleration."""
    W_vd, sdd = O.aba(model, env.p[None], env.q[None], env.s[None], env.W_v_WB[None, 0:3], env.W_v_WB[None, 3:6],
                      env.sd[None], tau[None], W_f_L[None])
    W_vd, sdd = W_vd[0], sdd[0]
    if not model.floating_base:
        return np.zeros(6, dtype=W_vd.dtype), sdd
    W_H_BW = env.W_H_B.copy()
    W_H_BW[0:3, 0:3] = np.eye(3)
    W_pd_B = env.base_velocity("mixed")[0:3]
    W_v_W_BW = np.zeros(6, dtype=W_vd.dtype)
    W_v_W_BW[0:3] = W_pd_B
    C_X_W = O.adjoint_from_transform(W_H_BW, inverse=True)
    return C_X_W @ (W_vd - O.cross_vx(W_v_W_BW) @ env.W_v_WB), sdd


def contact_problem(model, env: _Env, tau, W_f_L):
    """Everything ``compute_contact_forces`` builds before calling the QP solver
    (``rigid.py:268-352``): returns dict(Q, q, inactive, W_p_C, J, Minv)."""
    dt = env.s.dtype
    W_p_C, W_pd_C = O.collidable_points_pos_vel(model, env.W_H_L[None], env.W_v_WL[None])
    W_p_C, W_pd_C = W_p_C[0], W_pd_C[0]
    delta, delta_dot, n_hat = O.compute_penetration_data(model, W_p_C, W_pd_C)
    BW_nu = env.generalized_velocity("mixed")
    Minv = mass_matrix_inverse_mixed(model, env)
    J = contact_jacobian_mixed(model, env, "mixed")
    Jd = contact_jacobian_derivative_mixed(model, env)
    vd_free, sdd_free = aba_mixed(model, env, tau, W_f_L)
    BW_nud_free = np.concatenate([vd_free, sdd_free])
    nc = J.shape[0]
    a = (Jd.reshape(nc * 6, -1) @ BW_nu + J.reshape(nc * 6, -1) @ BW_nud_free).reshape(nc, 6)[:, 0:3]
    inactive = delta <= 0
    baum = np.where(inactive[:, None], 0.0, (dt.type(model.K) * delta + dt.type(model.D) * delta_dot)[:, None] * n_hat)
    Jl = J[:, 0:3, :].reshape(nc * 3, -1)
    delassus = Jl @ Minv @ Jl.T
    Q = delassus + model.regularization_delassus * np.eye(3 * nc)
    q = a.reshape(-1) - baum.reshape(-1)
    return dict(Q=Q, q=q, inactive=inactive, W_p_C=W_p_C, J=J, Minv=Minv, delassus=delassus, free_acc=a)


def compute_contact_forces(model, env: _Env, tau, W_f_L):
    """``RigidContacts.compute_contact_forces`` (``rigid.py:222-383``) -> W_f_C (nc,6) inertial-fixed."""
    pr = contact_problem(model, env, tau, W_f_L)
    x = _solve_contact_qp(pr["Q"], pr["q"], pr["inactive"], float(model.mu))
    CW_fl = x.reshape(-1, 3).astype(env.s.dtype)
    W_H_C = contact_transforms(model, env)
    f6 = np.zeros((CW_fl.shape[0], 6), dtype=env.s.dtype)
    f6[:, 0:3] = CW_fl
    return O.other_representation_to_inertial(f6, "mixed", W_H_C, is_force=True)


# ---------------------------------------------------------------------------------------
# rbda/contacts/relaxed_rigid.py
# ---------------------------------------------------------------------------------------


def relaxed_regularizers(model, link_idx, pos, vel):
    """``RelaxedRigidContacts._regularizers`` (``relaxed_rigid.py:533-653``) for the enabled
    points: ``pos`` (nc,3) position in the constraint frame (= -delta n), ``vel`` (nc,3).
    Returns (a_ref (3nc,), r (3nc,)).  As in the reference, the impedance is evaluated per
    COMPONENT of ``pos`` (``imp_x = |pos| / width`` is a 3-vector) and the parameters K / D are
    shadowed by the values derived from the time constant (``:585-591``)."""
    prm = model.relaxed
    Om, ze, xmin, xmax = prm["time_constant"], prm["damping_coefficient"], prm["d_min"], prm["d_max"]
    width, mid, pw = prm["width"], prm["midpoint"], prm["power"]
    mu = float(model.mu)
    mass = np.asarray(model.kin_dyn_parameters.link_parameters.mass, dtype=np.float64)
    imp_x = np.abs(pos) / width
    with np.errstate(invalid="ignore"):
        imp_a = (1.0 / np.power(mid, pw - 1)) * np.power(imp_x, pw)
        imp_b = 1 - (1.0 / np.power(1 - mid, pw - 1)) * np.power(1 - imp_x, pw)  # NaN for imp_x > 1 and odd powers: masked below
    imp_y = np.where(imp_x < mid, imp_a, imp_b)
    xi = xmin + imp_y * (xmax - xmin)
    xi = np.clip(xi, xmin, xmax)
    xi = np.where(imp_x > 1.0, xmax, xi)
    K = 1 / (xmax * Om * ze) ** 2
    D = 2 / (xmax * Om)
    a_ref = -(D * vel + K * xi * pos)
    # (vector) @ inv(M_L[link, :3, :3]) with M_L[:3, :3] = m 1  (:619-623)
    r = (2 * mu**2 * (1 - xi) / (xi + 1e-12)) * (1 + mu**2) / mass[link_idx][:, None]
    active = (np.einsum("ij,ij->i", pos, pos) > 0).astype(np.float64)[:, None]
    return (a_ref * active).reshape(-1), (r * active).reshape(-1)


def relaxed_contact_problem(model, env: _Env, tau, W_f_L):
    """What ``RelaxedRigidContacts.compute_contact_forces`` assembles before the solver
    (``relaxed_rigid.py:283-398``): A = J M^-1 J' + diag(r), b = J nu_dot_free + J_dot nu - a_ref,
    with the rows of inactive points zeroed."""
    W_p_C, W_pd_C = O.collidable_points_pos_vel(model, env.W_H_L[None], env.W_v_WL[None])
    W_p_C, W_pd_C = W_p_C[0], W_pd_C[0]
    en = _enabled(model)[0]
    W_p_C, W_pd_C = W_p_C[en], W_pd_C[en]
    delta, _, n_hat = O.compute_penetration_data(model, W_p_C, W_pd_C)
    pos = -delta[:, None] * n_hat
    body = np.asarray(model.kin_dyn_parameters.contact_parameters.body, dtype=int)[en]
    a_ref, r = relaxed_regularizers(model, body, pos, W_pd_C)
    BW_nu = env.generalized_velocity("mixed")
    vd_free, sdd_free = aba_mixed(model, env, tau, W_f_L)
    BW_nud_free = np.concatenate([vd_free, sdd_free])
    Minv = mass_matrix_inverse_mixed(model, env)
    act = (delta > 0)[:, None, None]
    Jl = (contact_jacobian_mixed(model, env, "mixed")[:, 0:3, :] * act).reshape(-1, BW_nu.shape[0])
    Jdl = (contact_jacobian_derivative_mixed(model, env)[:, 0:3, :] * act).reshape(-1, BW_nu.shape[0])
    A = Jl @ Minv @ Jl.T + np.diag(r)
    b = Jl @ BW_nud_free + Jdl @ BW_nu - a_ref
    return dict(A=A, b=b, active=delta > 0)


def relaxed_compute_contact_forces(model, env: _Env, tau, W_f_L):
    """``RelaxedRigidContacts.compute_contact_forces`` -> W_f_C (nc,6) inertial-fixed.  The
    reference minimises |A x + b|^2 with L-BFGS started from forces that are zero on the
    inactive points (``relaxed_rigid.py:465-505``); the minimiser is x = -A^-1 b on the active
    block (A is positive definite there: Delassus + positive diagonal) and zero elsewhere --
    parity is defined on that optimum, like for the rigid QP."""
    pr = relaxed_contact_problem(model, env, tau, W_f_L)
    A, b, act = pr["A"], pr["b"], pr["active"]
    x = np.zeros_like(b)
    sel = np.repeat(act, 3)
    if sel.any():
        x[sel] = -np.linalg.solve(A[np.ix_(sel, sel)], b[sel])
    CW_fl = x.reshape(-1, 3).astype(env.s.dtype)
    W_H_C = contact_transforms(model, env)
    f6 = np.zeros((CW_fl.shape[0], 6), dtype=env.s.dtype)
    f6[:, 0:3] = CW_fl
    return O.other_representation_to_inertial(f6, "mixed", W_H_C, is_force=True)


def compute_impact_velocity(inactive, M, J_WC, nu):
    """``RigidContacts.compute_impact_velocity`` (``rigid.py:163-220``)."""
    Jl = J_WC[:, 0:3, :].copy()
    Jl[inactive] = 0
    Jl = Jl.reshape(-1, J_WC.shape[-1])
    k = Jl.shape[0]
    A = np.block([[M, -Jl.T], [Jl, np.zeros((k, k), dtype=M.dtype)]])
    b = np.concatenate([M @ nu, np.zeros(k, dtype=M.dtype)])
    x = np.linalg.lstsq(A, b, rcond=None)[0]
    return x[0:M.shape[0]]


def update_velocity_after_impact(model, data):
    """``RigidContacts.update_velocity_after_impact`` (``rigid.py:385-436``), batched wrapper."""
    B = data.joint_positions.shape[0]
    v_lin = data.base_linear_velocity.copy()
    omega = data.base_angular_velocity.copy()
    sd = data.joint_velocities.copy()
    for b in range(B):
        env = _Env(data, b)
        idx2 = np.clip(_enabled(model)[0], 0, len(_enabled(model)[0]) - 1)  # second indexing, JAX clamps
        W_p_C, _ = O.collidable_points_pos_vel(model, env.W_H_L[None], env.W_v_WL[None])
        W_p_C = W_p_C[0][idx2]
        delta, _, _ = O.compute_penetration_data(model, W_p_C, np.zeros_like(W_p_C))
        J = contact_jacobian_mixed(model, env, "mixed")[idx2]
        M = mass_matrix_mixed(model, env)
        nu_pre = env.generalized_velocity("mixed")
        nu_post = compute_impact_velocity(delta <= 0, M, J, nu_pre)
        W_H_BW = env.W_H_B.copy()
        W_H_BW[0:3, 0:3] = np.eye(3)
        v6 = O.other_representation_to_inertial(nu_post[0:6], "mixed", W_H_BW, is_force=False)
        if model.floating_base:
            v_lin[b], omega[b] = v6[0:3], v6[3:6]
        sd[b] = nu_post[6:]
    out = type(data)(**{f: getattr(data, f) for f in data.__dataclass_fields__})
    out.base_linear_velocity = v_lin.astype(data.joint_positions.dtype)
    out.base_angular_velocity = omega.astype(data.joint_positions.dtype)
    out.joint_velocities = sd.astype(data.joint_positions.dtype)
    return out  # caches deliberately NOT refreshed (dataclasses.replace, rigid.py:429-434)


def link_contact_forces(model, data, W_f_L_external, tau_total):
    """``js.contact.link_contact_forces`` (``api/contact.py:514-554``) for RigidContacts."""
    B = data.joint_positions.shape[0]
    out = np.zeros_like(W_f_L_external)
    for b in range(B):
        fn = relaxed_compute_contact_forces if model.contact_model == "relaxed" else compute_contact_forces
        W_f_C = fn(model, _Env(data, b), tau_total[b], W_f_L_external[b])
        out[b] = O.link_forces_from_contact_forces(model, W_f_C[None])[0]
    return out


def step(model, data, link_forces_inertial=None, joint_force_references=None):
    """``js.model.step`` (``api/model.py:2601-2681``) with ``RigidContacts`` + SemiImplicitEuler."""
    B = data.joint_positions.shape[0]
    dtype = data.joint_positions.dtype
    nL, n = model.number_of_links(), model.dofs()
    W_f = np.zeros((B, nL, 6), dtype=dtype) if link_forces_inertial is None else np.asarray(link_forces_inertial, dtype=dtype)
    tau_ref = np.zeros((B, n), dtype=dtype) if joint_force_references is None else np.asarray(joint_force_references, dtype=dtype)
    tau_total = O.compute_resultant_torques(model, data.joint_positions, data.joint_velocities, tau_ref)
    nc = len(model.kin_dyn_parameters.contact_parameters.body)
    W_f_total = W_f
    if nc > 0:
        W_f_total = W_f + link_contact_forces(model, data, W_f, tau_total).astype(dtype)
    data_tf = _integrate(model, data, W_f_total, tau_total)
    if nc > 0 and model.contact_model != "relaxed":  # RelaxedRigid: no impact step (relaxed_rigid.py:262-281)
        data_tf = update_velocity_after_impact(model, data_tf)
    return data_tf


def _integrate(model, data, W_f_L_total, tau_total):
    """``semi_implicit_euler_integration`` (``api/integrators.py:14-88``) given the total link
    forces: reuses the soft oracle's integrator with the contact model switched off (the
    rigid model has no contact state, ``rigid.py:438-453``)."""
    import copy

    m2 = copy.copy(model)
    m2.contact_model = "none"
    out = O.semi_implicit_euler_integration(m2, data, W_f_L_total, tau_total)
    return out
