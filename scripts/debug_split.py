"""Diagnostic: split vs monolithic rigid cascade, per input class (airborne / standing / touching / sunk)."""
import sys, pathlib
import numpy as np
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import jaxsim_b200.api as js
from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams
from oracle import jaxsim_oracle as O
from oracle import rigid_oracle as R
from tests import helpers as H

dev = torch.device("cuda:0")
def mk():
    return H.build_model("ergocub_like", contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build(K=1e4, D=20.0))
om0 = H.oracle_model(mk())
B = 96
def inp(n, seed, mode):
    return O.random_model_data(om0, n, seed=seed, in_contact=mode)
a, b, c, d = inp(B // 4, 41, False), inp(B // 4, 42, "flat"), inp(B // 4, 43, True), inp(B // 4, 44, "flat")
pd_ = d.base_position.copy(); pd_[:, 2] -= 0.004
cat = lambda f: np.concatenate([getattr(a, f), getattr(b, f), getattr(c, f), getattr(d, f)], axis=0)
p = np.concatenate([a.base_position, b.base_position, c.base_position, pd_], axis=0)
od = O.data_replace(om0, cat("joint_positions"), cat("joint_velocities"), cat("base_quaternion"), cat("base_linear_velocity"), cat("base_angular_velocity"), p)
ref = R.step(om0, od)
for dtype in (torch.float64, torch.float32):
    outs = []
    for mono in (False, True, True, False):
        model = mk()
        if mono:
            model.set_options(rigid_mono=True)
        pdat = H.to_product(model, od, dtype, dev)
        out = js.model.step(model, pdat)
        torch.cuda.synchronize()
        outs.append(out)
    for name, (i, j) in (("split vs mono", (0, 1)), ("mono vs mono", (1, 2)), ("split vs split", (0, 3))):
        print(dtype, name)
        for leaf in ("_joint_positions", "_joint_velocities", "_base_linear_velocity", "_base_angular_velocity"):
            x, y = getattr(outs[i], leaf), getattr(outs[j], leaf)
            per = (x - y).abs().reshape(B, -1).max(dim=1).values.reshape(4, B // 4).max(dim=1).values
            print("   %-26s classes a/b/c/d max abs diff: %s" % (leaf, ["%.2e" % v for v in per.tolist()]))
    for name, i in (("split vs oracle", 0), ("mono vs oracle", 1)):
        for leaf, rl in (("_joint_velocities", "joint_velocities"), ("_base_linear_velocity", "base_linear_velocity")):
            x = getattr(outs[i], leaf).double().cpu().numpy(); y = getattr(ref, rl)
            per = np.abs(x - y).reshape(B, -1).max(axis=1).reshape(4, B // 4).max(axis=1)
            print("   %s %-24s a/b/c/d: %s" % (name, leaf, ["%.2e" % v for v in per.tolist()]))
