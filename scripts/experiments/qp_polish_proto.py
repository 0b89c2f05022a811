"""NumPy prototype: the CUDA interior-point method (qp_pyramids) + an active-set polish, on contact QPs of standing
ErgoCub-like environments built by the oracle."""
import sys, pathlib
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent.parent))
from tests import helpers as H
from oracle import jaxsim_oracle as O
from oracle import rigid_oracle as R
from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams

def Gmat(na, mu):
    G1 = np.array([[1, 0, -mu], [0, 1, -mu], [-1, 0, -mu], [0, -1, -mu], [0, 0, -1]], dtype=np.float64)
    G = np.zeros((5 * na, 3 * na))
    for a in range(na):
        G[5 * a:5 * a + 5, 3 * a:3 * a + 3] = G1
    return G

def polish(Q, q, G, s, z, mu_f, eps=1e-9):
    """active faces: z_j > s_j.  Returns (x, ok)."""
    na = Q.shape[0] // 3
    act = z > s
    cols = []
    for a in range(na):
        f = act[5 * a:5 * a + 5]
        if f[4] or (f[0] and f[2]) or (f[1] and f[3]):
            continue  # apex: x_a = 0
        sx = mu_f if f[0] else (-mu_f if f[2] else None)
        sy = mu_f if f[1] else (-mu_f if f[3] else None)
        e = np.zeros((3 * na,))
        def col(v):
            c = np.zeros(3 * na); c[3 * a:3 * a + 3] = v; cols.append(c)
        if sx is None and sy is None:
            col([1, 0, 0]); col([0, 1, 0]); col([0, 0, 1])
        elif sx is not None and sy is None:
            col([sx, 0, 1]); col([0, 1, 0])
        elif sx is None and sy is not None:
            col([0, sy, 1]); col([1, 0, 0])
        else:
            col([sx, sy, 1])
    N = 3 * na
    if cols:
        Z = np.array(cols).T
        y = np.linalg.solve(Z.T @ Q @ Z, -(Z.T @ q))
        x = Z @ y
    else:
        x = np.zeros(N)
    # checks
    r = Q @ x + q
    gx = G @ x
    scale_x = max(1.0, np.abs(x).max()); scale_r = max(1.0, np.abs(q).max())
    if (gx > eps * scale_x).any():
        return x, False
    # dual feasibility per point: -r_a = sum lambda_j n_j, lambda >= 0 over the ACTIVE faces (tight ones)
    for a in range(na):
        v = -r[3 * a:3 * a + 3]
        g = gx[5 * a:5 * a + 5]
        tight = g > -eps * scale_x
        # multipliers by non-negative least squares over tight faces (small: enumerate via lstsq + sign check)
        n = Gmat(1, mu_f)[tight]
        if n.shape[0] == 0:
            if np.abs(v).max() > 1e-7 * scale_r: return x, False
            continue
        from scipy.optimize import nnls
        lam, res = nnls(n.T, v)
        if res > 1e-7 * scale_r:
            return x, False
    return x, True

def ipm(Q, q, mu_f, tol, polish_from=None, max_iter=60):
    N = Q.shape[0]; na = N // 3; M = 5 * na
    G = Gmat(na, mu_f)
    x = np.zeros(N); s = np.ones(M); z = np.ones(M)
    qm = np.abs(q).max()
    best = 1e30; xb = x.copy()
    attempts = 0
    for it in range(max_iter):
        Qx = Q @ x
        rd = Qx + q + G.T @ z
        rp = G @ x + s
        mu = s @ z / M
        m_d = np.abs(rd).max() / (1 + qm + np.abs(Qx).max()); m_p = np.abs(rp).max() / (1 + np.abs(x).max())
        m_g = mu / (1 + abs(0.5 * x @ Qx + q @ x))
        merit = max(m_d, m_p, m_g)
        if merit < best: best = merit; xb = x.copy()
        if merit <= tol: return xb, it, attempts, "ipm"
        if polish_from is not None and merit < polish_from and attempts < 4:
            attempts += 1
            xp, ok = polish(Q, q, G, s, z, mu_f)
            if ok: return xp, it, attempts, "polish"
        W = z / s
        Hm = Q + G.T @ (W[:, None] * G)
        if not (m_g > 1e-3 * tol) or not (mu > 0): return xb, it, attempts, "gap"
        try:
            L = np.linalg.cholesky(Hm)
        except np.linalg.LinAlgError:
            return xb, it, attempts, "chol"
        sol = lambda b: np.linalg.solve(L.T, np.linalg.solve(L, b))
        dxa = sol(-(rd + G.T @ (z * (rp - s) / s)))
        dsa = -rp - G @ dxa; dza = -(s * z + z * dsa) / s
        rmax = max(1.0, (-dsa / s).max(), (-dza / z).max()); amax = 1 / rmax
        mua = ((s + amax * dsa) @ (z + amax * dza)) / M
        sg = (mua / mu) ** 3
        rc = s * z + dsa * dza - sg * mu
        dx = sol(-(rd + G.T @ ((z * rp - rc) / s)))
        ds = -rp - G @ dx; dz = -(rc + z * ds) / s
        rm2 = max(0.0, (-ds / s).max(), (-dz / z).max())
        al = 0.99 / rm2 if rm2 > 0.99 else 1.0
        x = x + al * dx; s = s + al * ds; z = z + al * dz
    return xb, max_iter, attempts, "maxit"

m = H.build_model("ergocub_like", contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build(K=1e4, D=20.0))
om = H.oracle_model(m)
B = 40
od = O.random_model_data(om, B, seed=3, in_contact="flat")
rng = np.random.default_rng(0)
tau = 10 * rng.uniform(size=(B, om.dofs()))
tot = {None: [], 1e-2: [], 1e-3: [], 1e-4: []}
err = []
for e in range(B):
    env = R._Env(od, e)
    pr = R.contact_problem(om, env, tau[e], np.zeros((om.number_of_links(), 6)))
    act = np.where(~pr["inactive"])[0]
    if len(act) == 0: continue
    sel = (3 * act[:, None] + np.arange(3)[None, :]).reshape(-1)
    Q = pr["Q"][np.ix_(sel, sel)]; q = pr["q"][sel]
    xr, it0, _, _ = ipm(Q, q, om.mu if hasattr(om, "mu") else 0.5, 1e-11)
    row = [len(act), it0]
    x8, it8, _, _ = ipm(Q, q, 0.5, 1e-8)
    row.append(it8)
    for pf in (1e-2, 1e-3, 1e-4):
        xp, itp, att, how = ipm(Q, q, 0.5, 1e-8, polish_from=pf)
        row += [itp, att, how, "%.1e" % (np.abs(xp - xr).max() / max(1, np.abs(xr).max()))]
    print(row)

print("---- scaling experiment: iterations to 1e-8 with q scaled by 1/sigma (x = sigma x')")
its = {k: [] for k in ("none", "qm", "qm/dmean", "qm/dmax", "qm/dmin")}
errs = {k: [] for k in its}
for e in range(B):
    env = R._Env(od, e)
    pr = R.contact_problem(om, env, tau[e], np.zeros((om.number_of_links(), 6)))
    act = np.where(~pr["inactive"])[0]
    if len(act) == 0: continue
    sel = (3 * act[:, None] + np.arange(3)[None, :]).reshape(-1)
    Q = pr["Q"][np.ix_(sel, sel)]; q = pr["q"][sel]
    xr, _, _, _ = ipm(Q, q, 0.5, 1e-11)
    qm = np.abs(q).max(); dg = np.diag(Q)
    for k, sig in (("none", 1.0), ("qm", qm), ("qm/dmean", qm / dg.mean()), ("qm/dmax", qm / dg.max()), ("qm/dmin", qm / dg.min())):
        x, it, _, how = ipm(Q, q / sig, 0.5, 1e-8)
        its[k].append(it); errs[k].append(np.abs(sig * x - xr).max() / max(1, np.abs(xr).max()))
for k in its:
    print(k, "mean it %.2f max %d" % (np.mean(its[k]), np.max(its[k])), "median err %.1e" % np.median(errs[k]))
print("qm range", [float("%.3g" % v) for v in (min(np.abs(R.contact_problem(om, R._Env(od, e), tau[e], np.zeros((om.number_of_links(), 6)))["q"]).max() for e in range(5)),)])

print("---- sigma = c * qm, tol 1e-8 and 1e-11; unscaled relative merit of the answer re-evaluated")
def rel_merit(Q, q, x):
    # KKT residual of x for the cone-constrained QP, via non-negative least squares of the multipliers
    from scipy.optimize import nnls
    na = Q.shape[0] // 3
    G = Gmat(na, 0.5)
    r = Q @ x + q
    lam, res = nnls(G.T, -r)
    return res / (1 + np.abs(q).max()), max(0.0, (G @ x).max()) / (1 + np.abs(x).max())
for tol in (1e-8, 1e-11):
    for c in (None, 0.1, 1.0, 10.0, 100.0):
        its, md, mp, bad = [], [], [], 0
        for e in range(B):
            env = R._Env(od, e)
            pr = R.contact_problem(om, env, tau[e], np.zeros((om.number_of_links(), 6)))
            act = np.where(~pr["inactive"])[0]
            if len(act) == 0: continue
            sel = (3 * act[:, None] + np.arange(3)[None, :]).reshape(-1)
            Q = pr["Q"][np.ix_(sel, sel)]; q = pr["q"][sel]
            sig = 1.0 if c is None else c * np.abs(q).max()
            x, it, _, how = ipm(Q, q / sig, 0.5, tol)
            a, b_ = rel_merit(Q, q, sig * x)
            its.append(it); md.append(a); mp.append(b_); bad += how not in ("ipm",)
        print("tol %.0e c %s: mean it %.2f max %d | dual residual median %.1e max %.1e | primal max %.1e | non-converged exits %d" % (tol, c, np.mean(its), np.max(its), np.median(md), np.max(md), np.max(mp), bad))

print("---- sanity: scaled (c = 1) vs unscaled answers, compared through Q x (the well-conditioned quantity) and x itself")
dQ, dX, nX = [], [], []
for e in range(12):
    env = R._Env(od, e)
    pr = R.contact_problem(om, env, tau[e], np.zeros((om.number_of_links(), 6)))
    act = np.where(~pr["inactive"])[0]
    if len(act) == 0: continue
    sel = (3 * act[:, None] + np.arange(3)[None, :]).reshape(-1)
    Q = pr["Q"][np.ix_(sel, sel)]; q = pr["q"][sel]
    xr, _, _, _ = ipm(Q, q, 0.5, 1e-11)
    sig = np.abs(q).max()
    xs, it, _, _ = ipm(Q, q / sig, 0.5, 1e-8); xs = sig * xs
    xu, itu, _, _ = ipm(Q, q, 0.5, 1e-8)
    dQ.append((np.abs(Q @ (xs - xr)).max() / np.abs(Q @ xr).max(), np.abs(Q @ (xu - xr)).max() / np.abs(Q @ xr).max()))
    dX.append((np.abs(xs - xr).max() / np.abs(xr).max(), np.abs(xu - xr).max() / np.abs(xr).max()))
    nX.append(np.abs(xr).max())
print("Qx rel diff (scaled, unscaled) vs 1e-11 answer:", ["%.1e/%.1e" % t for t in dQ])
print("x  rel diff (scaled, unscaled):", ["%.1e/%.1e" % t for t in dX])
print("|x|max", ["%.0f" % v for v in nX])

print("---- accuracy-matched comparison: geometric-mean rel. error of Q x vs the 1e-12 answer, and iterations")
def run(scale_c, tol):
    its, errs = [], []
    for e in range(B):
        env = R._Env(od, e)
        pr = R.contact_problem(om, env, tau[e], np.zeros((om.number_of_links(), 6)))
        act = np.where(~pr["inactive"])[0]
        if len(act) == 0: continue
        sel = (3 * act[:, None] + np.arange(3)[None, :]).reshape(-1)
        Q = pr["Q"][np.ix_(sel, sel)]; q = pr["q"][sel]
        if e not in REF:
            REF[e] = ipm(Q, q / np.abs(q).max(), 0.5, 1e-13)[0] * np.abs(q).max()
        xr = REF[e]
        sig = 1.0 if scale_c is None else scale_c * np.abs(q).max()
        x, it, _, _ = ipm(Q, q / sig, 0.5, tol)
        its.append(it); errs.append(max(np.abs(Q @ (sig * x - xr)).max() / np.abs(Q @ xr).max(), 1e-16))
    return np.mean(its), np.exp(np.mean(np.log(errs))), np.max(errs)
REF = {}
for c, tol in ((None, 1e-8), (1.0, 1e-8), (1.0, 1e-9), (1.0, 1e-10), (0.1, 1e-9), (0.1, 1e-10), (None, 1e-11), (1.0, 1e-11), (1.0, 1e-12), (1.0, 1e-13)):
    print("c %s tol %.0e: mean it %.2f | Qx err gmean %.1e max %.1e" % ((c, tol) + run(c, tol)))

print("---- two-parameter normalisation: x = alpha*sigma*y, Q' = alpha^2 Q (unit mean / max diagonal), q' = alpha q / sigma (unit max)")
def run2(which, tol, c=1.0):
    its, errs = [], []
    for e in range(B):
        env = R._Env(od, e)
        pr = R.contact_problem(om, env, tau[e], np.zeros((om.number_of_links(), 6)))
        act = np.where(~pr["inactive"])[0]
        if len(act) == 0: continue
        sel = (3 * act[:, None] + np.arange(3)[None, :]).reshape(-1)
        Q = pr["Q"][np.ix_(sel, sel)]; q = pr["q"][sel]
        xr = REF[e]
        d = np.diag(Q)
        alpha = 1 / np.sqrt({"mean": d.mean(), "max": d.max(), "min": d.min()}[which])
        sig = c * np.abs(alpha * q).max()
        y, it, _, _ = ipm(alpha * alpha * Q, alpha * q / sig, 0.5, tol)
        x = alpha * sig * y
        its.append(it); errs.append(max(np.abs(Q @ (x - xr)).max() / np.abs(Q @ xr).max(), 1e-16))
    return np.mean(its), np.exp(np.mean(np.log(errs))), np.max(errs)
for which in ("mean", "max"):
    for c in (1.0, 0.3):
        for tol in (1e-8, 1e-9, 1e-10, 1e-11, 1e-12):
            print("diag %s c %.1f tol %.0e: mean it %.2f | Qx err gmean %.1e max %.1e" % ((which, c, tol) + run2(which, tol, c)))
