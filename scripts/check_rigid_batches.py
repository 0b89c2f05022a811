"""Split vs monolithic rigid cascade at odd batch sizes (1, 5, 33, 1000): bit-identical leaves, finite results."""
import pathlib, sys
import numpy as np
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import jaxsim_b200.api as js
from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams
from oracle import jaxsim_oracle as O
from tests import helpers as H

dev = torch.device("cuda:0")
mk = lambda: H.build_model("icub_like", contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build(K=1e4, D=20.0))
om = H.oracle_model(mk())
for B in (1, 5, 33, 1000):
    od = O.random_model_data(om, B, seed=B, in_contact="flat" if B % 2 else True)
    for dtype in (torch.float32, torch.float64):
        outs = []
        for mono in (False, True):
            m = mk()
            if mono:
                m.set_options(rigid_mono=True)
            d = H.to_product(m, od, dtype, dev)
            o = js.model.step(m, d)
            o = js.model.step_n(m, o, 3) if B <= 33 else js.model.step(m, o)   # in-place cascades from the second step on
            torch.cuda.synchronize()
            outs.append(o)
        for leaf in ("_joint_positions", "_joint_velocities", "_base_position", "_base_linear_velocity", "_link_velocities"):
            x, y = getattr(outs[0], leaf), getattr(outs[1], leaf)
            assert torch.isfinite(x).all(), (B, dtype, leaf)
            assert torch.equal(x, y), (B, dtype, leaf, float((x - y).abs().max()))
    print("B =", B, "ok")
