"""Run the UNMODIFIED reference step (over oracle/refshim) next to the NumPy oracle on the same inputs:

    PYTHONPATH=. python scripts/ref_try.py pendulum box icub_like        (build container only)
"""
import numpy as np, time, sys
from tests.golden import refenv
jaxsim, js = refenv.load()
from jaxsim_b200 import models
from tests import helpers as H
from oracle import jaxsim_oracle as O
for name in sys.argv[1:]:
    rm = refenv.reference_model(models.urdf(name))
    pm = H.build_model(name); om = H.oracle_model(pm)
    kd = rm.kin_dyn_parameters
    print(name, "parents equal:", np.array_equal(np.asarray(kd.parent_array), np.asarray(pm.kin_dyn_parameters.parent_array)))
    od = O.random_model_data(om, 2, seed=0, in_contact=True)
    ref = O.step(om, od)
    for e in range(2):
        t0 = time.time()
        d = js.data.JaxSimModelData.build(model=rm, base_position=od.base_position[e], base_quaternion=od.base_quaternion[e],
            joint_positions=od.joint_positions[e], joint_velocities=od.joint_velocities[e],
            base_linear_velocity=od.base_linear_velocity[e], base_angular_velocity=od.base_angular_velocity[e],
            velocity_representation=jaxsim.VelRepr.Inertial)
        out = js.model.step(model=rm, data=d)
        dt = time.time() - t0
        worst = 0
        for oname, pname in H.LEAVES:
            a = np.asarray(getattr(out, pname)); b = getattr(ref, oname)[e]
            err = np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-12) if b.size else 0
            worst = max(worst, err)
            if err > 1e-10: print("   ", oname, err)
        a = np.asarray(out.contact_state["tangential_deformation"]); b = ref.tangential_deformation[e]
        print("  env", e, "%.2fs" % dt, "worst rel err %.2e" % worst, "m err %.2e" % (np.max(np.abs(a-b)) if b.size else 0))
