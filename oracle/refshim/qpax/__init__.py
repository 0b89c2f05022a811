"""Stand-in for `qpax.solve_qp` (test infrastructure only, see ../README.md).

    min_x 0.5 x^T Q x + q^T x   s.t.  A x = b,  G x <= h

Primal-dual interior-point method with Mehrotra's predictor-corrector, written from the
published algorithm (Mattingley & Boyd, "CVXGEN", 2012, section 5 -- the one qpax cites).
The stand-in iterates to a tight tolerance regardless of `solver_tol` (the strictly convex
QP of the rigid contact model has a unique optimum), so the fixtures made with it hold the
exact optimum and the caller's `solver_tol` only shows up as the tolerance of a comparison.

The reference's contact QP pins the force of every inactive point to zero with pairs of
opposite inequalities (`f_z <= 0` and `-f_z <= 0`, rbda/contacts/rigid.py:472-497), i.e. the
feasible set has no interior and a tight-tolerance interior-point iteration cannot converge
on it.  Such implied equalities (rows g and -g with zero bound and a single free variable)
are detected and the pinned variables eliminated exactly before iterating -- the optimum is
unchanged.  Equality constraints are supported only when absent (A has zero rows), which is
how the reference calls it (rigid.py:349-350).
"""
import numpy as np

from jax._core import asarray as _arr


def _pinned_variables(G, h):
    """Variables forced to zero by pairs of opposite single-variable rows with zero bound."""
    n = G.shape[1]
    pinned = np.zeros(n, dtype=bool)
    changed = True
    while changed:
        changed = False
        Gf = np.where(pinned[None, :], 0.0, G)
        nz = Gf != 0
        single = np.where((nz.sum(axis=1) == 1) & (h == 0))[0]
        sign = {}
        for r in single:
            v = int(np.argmax(nz[r]))
            sign.setdefault(v, set()).add(np.sign(Gf[r, v]))
        for v, sg in sign.items():
            if not pinned[v] and {1.0, -1.0} <= sg:
                pinned[v] = True
                changed = True
    return pinned


def _pdip(Q, q, G, h, tol, max_iter):
    n, m = Q.shape[0], G.shape[0]
    x = np.zeros(n)
    if m == 0:
        return np.linalg.solve(Q, -q), np.zeros(0), np.zeros(0), True, 0
    s, z = np.ones(m), np.ones(m)
    best = (np.inf, x, s, z)
    it = 0
    for it in range(1, max_iter + 1):
        Qx = Q @ x
        r_d = Qx + q + G.T @ z
        r_p = G @ x + s - h
        mu = float(s @ z) / m
        m_d = np.abs(r_d).max() / (1.0 + np.abs(q).max() + np.abs(Qx).max())
        m_p = np.abs(r_p).max() / (1.0 + np.abs(x).max())
        m_g = mu / (1.0 + abs(0.5 * x @ Qx + q @ x))
        merit = max(m_d, m_p, m_g)
        if merit < best[0]:
            best = (merit, x.copy(), s.copy(), z.copy())
        if not (merit > tol) or not (m_g > 1e-3 * tol) or not np.isfinite(merit):
            break
        w = z / s
        try:
            L = np.linalg.cholesky(Q + G.T @ (w[:, None] * G))
        except np.linalg.LinAlgError:
            break

        def newton(r_c):
            rhs = -(r_d + G.T @ ((z * r_p - r_c) / s))
            dx = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
            ds = -r_p - G @ dx
            return dx, ds, -(r_c + z * ds) / s

        def max_step(v, dv):
            neg = dv < 0
            return float(np.min(-v[neg] / dv[neg])) if neg.any() else np.inf

        _, dsa, dza = newton(s * z)
        a_aff = min(1.0, max_step(s, dsa), max_step(z, dza))
        sigma = (float((s + a_aff * dsa) @ (z + a_aff * dza)) / m / mu) ** 3
        dx, ds, dz = newton(s * z + dsa * dza - sigma * mu)
        a = min(1.0, 0.99 * min(max_step(s, ds), max_step(z, dz)))
        x, s, z = x + a * dx, s + a * ds, z + a * dz
    merit, x, s, z = best
    return x, s, z, bool(merit < 1e-8), it


def solve_qp(Q, q, A, b, G, h, solver_tol=1e-3, max_iter=200, **_):
    Q, q, A, b, G, h = (np.asarray(v, dtype=float) for v in (Q, q, A, b, G, h))
    if A.shape[0] != 0:
        raise NotImplementedError("the qpax stand-in supports inequality-only problems")
    n, m = Q.shape[0], G.shape[0]
    pinned = _pinned_variables(G, h)
    free = ~pinned
    Gf = G[:, free]
    rows = np.abs(Gf).sum(axis=1) > 0
    x = np.zeros(n)
    s, z = np.zeros(m), np.zeros(m)
    conv, it = True, 0
    if free.any():
        xf, sf, zf, conv, it = _pdip(Q[np.ix_(free, free)], q[free], Gf[rows], h[rows], min(float(solver_tol), 1e-11), max_iter)
        x[free] = xf
        s[rows], z[rows] = sf, zf
    return _arr(x), _arr(s), _arr(z), _arr(np.zeros(0)), _arr(conv), _arr(it)
