import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import ctypes as C
import numpy as np, torch
from tests import helpers as H
from oracle import jaxsim_oracle as O, rigid_oracle as R
import jaxsim_b200.api as js
from jaxsim_b200 import _lib
from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams
dev = torch.device("cuda:0")
z = np.load("scratch/dbg_rigid.npz")
for K, D in ((1e4, 20.0), (0.0, 0.0)):
    model = H.build_model("ergocub_like", contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build(K=K, D=D))
    om = H.oracle_model(model)
    od = O.data_replace(om, z["joint_positions"], z["joint_velocities"], z["base_quaternion"], z["base_linear_velocity"],
                        z["base_angular_velocity"], z["base_position"])
    lib = _lib.load()
    for td in (torch.float64,):
        for use_tau in (True, False):
            pd = H.to_product(model, od, td, dev)
            tt = torch.as_tensor(z["tau"], dtype=td, device=dev) if use_tau else None
            dm = model.device_model(dev)
            cnt = (C.c_ulonglong * 8)()
            lib.b200sim_debug_counters(dm.handle, cnt)
            res = []
            for e in range(10):
                sub = js.data._map_leaves(pd, lambda t: t[e:e+1].contiguous())
                out = js.model.step(model, sub, joint_force_references=None if tt is None else tt[e:e+1])
                torch.cuda.synchronize()
                lib.b200sim_debug_counters(dm.handle, cnt)
                res.append((bool(torch.isfinite(out._joint_velocities).all()), list(cnt)[:8]))
            print("K", K, "tau", use_tau, td)
            for e, r in enumerate(res):
                print("   env", e, "finite", r[0], "[qp iters, qps, max it, act(QP), full items, impact-only, impacts, act(impact)]", r[1])
    ref = R.step(om, od, joint_force_references=z["tau"])
    print("oracle finite:", np.isfinite(ref.joint_velocities).all(axis=1))
