import sys, pathlib, os; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
os.environ["B200SIM_LIB"] = str(pathlib.Path(__file__).resolve().parents[1] / "scratch/libb200sim_dbg.so")
import ctypes as C
import numpy as np, torch
from tests import helpers as H
from oracle import jaxsim_oracle as O, rigid_oracle as R
import jaxsim_b200.api as js
from jaxsim_b200 import _lib
from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams
np.set_printoptions(precision=5, linewidth=220)
dev = torch.device("cuda:0")
z = np.load("scratch/dbg_rigid.npz")
model = H.build_model("ergocub_like", contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build(K=1e4, D=20.0))
om = H.oracle_model(model)
od = O.data_replace(om, z["joint_positions"], z["joint_velocities"], z["base_quaternion"], z["base_linear_velocity"],
                    z["base_angular_velocity"], z["base_position"])
lib = _lib.load()
td = torch.float64
pd = H.to_product(model, od, td, dev)
tt = torch.as_tensor(z["tau"], dtype=td, device=dev)
dm = model.device_model(dev)
cnt = (C.c_ulonglong * 8)()
lib.b200sim_debug_counters(dm.handle, cnt)
dump = (C.c_double * 512)()
for e in (1, 6):
    sub = js.data._map_leaves(pd, lambda t: t[e:e+1].contiguous())
    out = js.model.step(model, sub, joint_force_references=tt[e:e+1])
    torch.cuda.synchronize()
    lib.b200sim_debug_rigid_dump(dm.handle, dump)
    D = np.array(dump[:])
    na = int(D[0]); N = 3 * na
    print("env", e, "finite:", bool(torch.isfinite(out._joint_velocities).all()), "na", na, "qp_it", D[7])
    print("  a0 free:", D[1:7])
    print("  q:", D[8:8+N])
    print("  diag(Q):", D[104:104+N])
    print("  x:", D[200:200+N])
    print("  a0 total:", D[296:302])
    print("  sdd:", D[309:309+49])
    for leaf in ("_joint_positions", "_joint_velocities", "_base_quaternion", "_base_linear_velocity", "_base_position"):
        print("  ", leaf, "finite" if bool(torch.isfinite(getattr(out, leaf)).all()) else "NON-FINITE")
