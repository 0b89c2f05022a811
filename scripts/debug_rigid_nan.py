import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np, torch
from tests import helpers as H
from oracle import jaxsim_oracle as O, rigid_oracle as R
import jaxsim_b200.api as js
from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams
dev = torch.device("cuda:0")
model = H.build_model("ergocub_like", contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build(K=1e4, D=20.0))
om = H.oracle_model(model)
B = 16384
third = B // 3
parts = [O.random_model_data(om, third, seed=41, in_contact=False), O.random_model_data(om, third, seed=42, in_contact=True),
         O.random_model_data(om, B - 2 * third, seed=43, in_contact="flat")]
cat = lambda f: np.concatenate([getattr(p, f) for p in parts], axis=0)
perm = np.random.default_rng(3).permutation(B)
kind = (perm // third).clip(0, 2)
od = O.data_replace(om, cat("joint_positions")[perm], cat("joint_velocities")[perm], cat("base_quaternion")[perm],
                    cat("base_linear_velocity")[perm], cat("base_angular_velocity")[perm], cat("base_position")[perm])
tau = 10 * np.random.default_rng(2).uniform(size=(B, om.dofs())).astype(np.float32).astype(np.float64)
W_p_C, _ = O.collidable_points_pos_vel(om, od.link_transforms, od.link_velocities)
nact = (W_p_C[..., 2] < 0).sum(axis=1)
for td in (torch.float64, torch.float32):
    pd = H.to_product(model, od, td, dev)
    tt = torch.as_tensor(tau, dtype=td, device=dev)
    out = js.model.step(model, pd, joint_force_references=tt)
    torch.cuda.synchronize()
    bad = (~torch.isfinite(out._joint_velocities).all(dim=1)).nonzero().flatten().cpu().numpy()
    print(td, "non-finite envs:", len(bad), "kinds:", np.bincount(kind[bad], minlength=3), "active points of bad:", np.bincount(nact[bad], minlength=33)[:33])
    print("   active-point histogram of the batch:", np.bincount(nact, minlength=33)[:33])
    for b in bad[:5]:
        sub = O.data_replace(om, od.joint_positions[b:b+1], od.joint_velocities[b:b+1], od.base_quaternion[b:b+1],
                             od.base_linear_velocity[b:b+1], od.base_angular_velocity[b:b+1], od.base_position[b:b+1])
        one = js.model.step(model, H.to_product(model, sub, td, dev), joint_force_references=tt[b:b+1])
        print("   env", b, "kind", kind[b], "nact", nact[b], "alone finite:", bool(torch.isfinite(one._joint_velocities).all()),
              "min z", float(W_p_C[b, :, 2].min()))
    # sizes: does it depend on the batch?
    for Bq in (1024, 4096, 8192):
        subq = O.data_replace(om, od.joint_positions[:Bq], od.joint_velocities[:Bq], od.base_quaternion[:Bq],
                              od.base_linear_velocity[:Bq], od.base_angular_velocity[:Bq], od.base_position[:Bq])
        o2 = js.model.step(model, H.to_product(model, subq, td, dev), joint_force_references=tt[:Bq])
        print("   batch", Bq, "non-finite:", int((~torch.isfinite(o2._joint_velocities).all(dim=1)).sum()))
