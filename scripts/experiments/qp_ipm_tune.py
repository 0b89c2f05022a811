"""Tuning experiments for the normalised contact-QP interior-point method (NumPy, CPU)."""
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

def ipm(Q, q, mu_f, tol, eta=0.99, start="ones", sig_exp=3, max_iter=60, extra_corr=False):
    N = Q.shape[0]; na = N // 3; M = 5 * na
    G = Gmat(na, mu_f)
    x = np.zeros(N); s = np.ones(M); z = np.ones(M)
    if start == "mehrotra":
        x = np.linalg.solve(Q + G.T @ G, -q)
        s = -G @ x; z = s.copy()
        ds = max(-1.5 * s.min(), 0.0); dz = max(-1.5 * z.min(), 0.0)
        s = s + ds; z = z + dz
        sz = s @ z
        if sz > 0:
            s = s + 0.5 * sz / z.sum(); z = z + 0.5 * sz / s.sum()
        s = np.maximum(s, 1e-3); z = np.maximum(z, 1e-3)
    qm = np.abs(q).max()
    best = 1e30; xb = x.copy()
    for it in range(max_iter):
        Qx = Q @ x
        rd = Qx + q + G.T @ z; rp = G @ x + s; mu = s @ z / M
        merit = max(np.abs(rd).max() / (1 + qm + np.abs(Qx).max()), np.abs(rp).max() / (1 + np.abs(x).max()), mu / (1 + abs(0.5 * x @ Qx + q @ x)))
        if merit < best: best = merit; xb = x.copy()
        if merit <= tol: return xb, it
        W = z / s
        try:
            L = np.linalg.cholesky(Q + G.T @ (W[:, None] * G))
        except np.linalg.LinAlgError:
            return xb, it
        sol = lambda b: np.linalg.solve(L.T, np.linalg.solve(L, b))
        dxa = sol(-(rd + G.T @ (z * (rp - s) / s)))
        dsa = -rp - G @ dxa; dza = -(s * z + z * dsa) / s
        amax = 1 / max(1.0, (-dsa / s).max(), (-dza / z).max())
        mua = ((s + amax * dsa) @ (z + amax * dza)) / M
        sg = (mua / mu) ** sig_exp
        rc = s * z + dsa * dza - sg * mu
        dx = sol(-(rd + G.T @ ((z * rp - rc) / s)))
        ds = -rp - G @ dx; dz = -(rc + z * ds) / s
        rm2 = max(0.0, (-ds / s).max(), (-dz / z).max())
        if eta == "adaptive":
            e = max(0.99, 1 - mu)  # closer to the boundary as the gap closes
        else:
            e = eta
        al = e / rm2 if rm2 > e else 1.0
        x = x + al * dx; s = s + al * ds; z = z + al * dz
    return xb, max_iter

m = H.build_model("ergocub_like", contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build(K=1e4, D=20.0))
om = H.oracle_model(m)
B = 40
od = O.random_model_data(om, B, seed=3, in_contact="flat")
tau = 10 * np.random.default_rng(0).uniform(size=(B, om.dofs()))
probs = []
for e in range(B):
    pr = R.contact_problem(om, R._Env(od, e), tau[e], np.zeros((om.number_of_links(), 6)))
    act = np.where(~pr["inactive"])[0]
    if len(act) == 0: continue
    sel = (3 * act[:, None] + np.arange(3)[None, :]).reshape(-1)
    Q = pr["Q"][np.ix_(sel, sel)]; q = pr["q"][sel]
    a2 = len(sel) / np.trace(Q); al = np.sqrt(a2)
    probs.append((Q, q, a2, al))
refs = []
for Q, q, a2, al in probs:
    sg = 0.3 * np.abs(al * q).max()
    refs.append(al * sg * ipm(a2 * Q, al * q / sg, 0.5, 1e-14)[0])
def run(c=0.3, tol=1e-9, **kw):
    its, errs = [], []
    for (Q, q, a2, al), xr in zip(probs, refs):
        sg = c * np.abs(al * q).max()
        y, it = ipm(a2 * Q, al * q / sg, 0.5, tol, **kw)
        x = al * sg * y
        its.append(it); errs.append(max(np.abs(Q @ (x - xr)).max() / np.abs(Q @ xr).max(), 1e-16))
    return "mean it %.2f max %d | Qx err gmean %.1e max %.1e" % (np.mean(its), np.max(its), np.exp(np.mean(np.log(errs))), np.max(errs))
print("baseline c=0.3 eta=.99        ", run())
for c in (0.1, 1.0, 3.0):
    print("c=%.1f                        " % c, run(c=c))
for eta in (0.995, 0.999, "adaptive"):
    print("eta=%s                   " % eta, run(eta=eta))
print("mehrotra start                ", run(start="mehrotra"))
print("mehrotra start c=1            ", run(c=1.0, start="mehrotra"))
for se in (2, 4):
    print("sigma exponent %d             " % se, run(sig_exp=se))
print("eta .999 + tol 1e-9 c=0.1     ", run(c=0.1, eta=0.999))
