"""One small RigidContacts step through the split cascade (every route: airborne, level 1, level 2, landing) for
compute-sanitizer: checked against the oracle."""
import pathlib, sys
import numpy as np
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import jaxsim_b200.api as js
from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams
from oracle import jaxsim_oracle as O
from oracle import rigid_oracle as R
from tests import helpers as H

dev = torch.device("cuda:0")
model = H.build_model("ergocub_like", contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build(K=1e4, D=20.0))
om = H.oracle_model(model)
a, b, c, d = (O.random_model_data(om, 4, seed=s, in_contact=m) for s, m in ((41, False), (42, "flat"), (43, True), (44, "flat")))
pd_ = d.base_position.copy(); pd_[:, 2] -= 0.004
cat = lambda f: np.concatenate([getattr(a, f), getattr(b, f), getattr(c, f), getattr(d, f)], axis=0)
od = O.data_replace(om, cat("joint_positions"), cat("joint_velocities"), cat("base_quaternion"), cat("base_linear_velocity"),
                    cat("base_angular_velocity"), np.concatenate([a.base_position, b.base_position, c.base_position, pd_], axis=0))
ref = R.step(om, od)
for dtype, key in ((torch.float64, "float64"), (torch.float32, "float32")):
    out = js.model.step(model, H.to_product(model, od, dtype, dev))
    torch.cuda.synchronize()
    v = max(float(np.abs(od.joint_velocities).max()), float(np.abs(od.base_linear_velocity).max()), 1e-3)
    errs = H.compare_data(out, ref, H.RTOL[key], f"rigid split {key}",
                          floors={"base_linear_velocity": v, "base_angular_velocity": v, "joint_velocities": v, "link_velocities": v})
    print(f"rigid split cascade {key}: max rel err {max(errs.values()):.3e}")
