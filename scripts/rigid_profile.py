"""Drive the rigid-contact step for profiling: ErgoCub-like model, 'standing' inputs (level
base, soles 2-5 mm into the ground), a few eager steps bracketed by cudaProfilerStart/Stop.

    ncu --profile-from-start off --set full --import-source on -k regex:rigid_step_kernel \
        -o gpurun_out/rigid python scripts/rigid_profile.py --batch 4096
"""
import argparse
import pathlib
import sys

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import jaxsim_b200.api as js  # noqa: E402
from jaxsim_b200 import models  # noqa: E402
from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--model", default="ergocub_like")
ap.add_argument("--dtype", default="f32")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--qp32", action="store_true")
ap.add_argument("--mono", action="store_true", help="contact QP inside the rigid kernel (B200SIM_OPT_RIGID_MONO)")
ap.add_argument("--inputs", default="standing", choices=("standing", "random"))
args = ap.parse_args()
dev = torch.device("cuda:0")
dtype = torch.float32 if args.dtype == "f32" else torch.float64
m = js.model.JaxSimModel.build_from_model_description(
    models.urdf(args.model), time_step=1e-3, contact_model=RigidContacts.build(),
    contact_params=RigidContactsParams.build(K=1e4, D=20.0))
if args.qp32 or args.mono:
    m.set_options(rigid_qp_f32=args.qp32, rigid_mono=args.mono)
B, n = args.batch, m.dofs()
gen = torch.Generator(device=dev).manual_seed(0)
u = lambda *sh: 2 * torch.rand(*sh, dtype=dtype, device=dev, generator=gen) - 1  # noqa: E731
rpy = 1e-3 * u(B, 3)
q = torch.cat([torch.ones(B, 1, dtype=dtype, device=dev), 0.5 * rpy], dim=-1)
q = q / q.norm(dim=-1, keepdim=True)
p = torch.cat([u(B, 2), torch.ones(B, 1, dtype=dtype, device=dev)], dim=-1)
kw = dict(base_quaternion=q, joint_positions=1e-3 * u(B, n), joint_velocities=0.1 * u(B, n),
          base_linear_velocity=0.1 * u(B, 3), base_angular_velocity=0.1 * u(B, 3),
          velocity_representation=js.common.VelRepr.Inertial, batch_size=B, dtype=dtype, device=dev)
d0 = js.data.JaxSimModelData.build(m, base_position=p, **kw)
cp = m.kin_dyn_parameters.contact_parameters
body = torch.as_tensor(np.array(cp.body), device=dev)
Lp = torch.as_tensor(np.asarray(cp.point), dtype=dtype, device=dev)
H = d0.link_transforms[:, body]
z = (H[..., 2, 0:3] * Lp).sum(-1) + H[..., 2, 3]
p[:, 2] -= z.min(dim=1).values + 0.002 + 0.003 * torch.rand(B, dtype=dtype, device=dev, generator=gen)
data = js.data.JaxSimModelData.build(m, base_position=p, **kw)
if args.inputs == "random":
    data = js.data.random_model_data(m, batch_size=B, seed=51, dtype=dtype, device=dev, velocity_representation=js.common.VelRepr.Inertial)
tau = 10 * torch.rand(B, n, dtype=dtype, device=dev)
out = js.model.step(m, data, joint_force_references=tau)
import ctypes  # noqa: E402
from jaxsim_b200 import _lib  # noqa: E402
lib = _lib.load()
cnt = (ctypes.c_ulonglong * 8)()
lib.b200sim_debug_counters.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
lib.b200sim_debug_counters(m.device_model(dev).handle, cnt)  # enable
js.model.step(m, data, joint_force_references=tau, out=out)
lib.b200sim_debug_counters(m.device_model(dev).handle, cnt)
c = list(cnt)
print("counters: QPs %d, mean it %.1f, max it %d, mean active %.1f | full items %d, impact-only %d, impacts %d (mean active %.1f)"
      % (c[1], c[0] / max(c[1], 1), c[2], c[3] / max(c[1], 1), c[4], c[5], c[6], c[7] / max(c[6], 1)))
for _ in range(3):
    js.model.step(m, data, joint_force_references=tau, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
for _ in range(args.steps):
    js.model.step(m, data, joint_force_references=tau, out=out)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ms = e0.elapsed_time(e1) / args.steps
clk = (ctypes.c_ulonglong * 32)()
lib.b200sim_debug_phase_clocks.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
lib.b200sim_debug_phase_clocks(m.device_model(dev).handle, clk)
NAMES = {1: "start of item", 2: "kinematics + contact points", 3: "link init + ABA passes 2, 3 (free acceleration)", 4: "Delassus (contact)",
         5: "QP / linear solve", 6: "forces -> tree response", 7: "semi-implicit Euler", 8: "kinematics of the new state + cache stores",
         9: "contact points (impact)", 10: "link init + pass 2 + Delassus (impact)", 11: "pivoted Cholesky (impact)", 12: "impulse response + stores"}
t = [int(v) for v in clk][16:29]
prev = t[1]
print("timeline of warp 0 / block 0, first work item of the LAST rigid-kernel launch (clock64, us at 1965 MHz):")
for k in range(2, 13):
    if t[k] == 0 or t[k] < prev:
        continue
    print("  %-52s %9d cyc %8.1f us" % (NAMES[k], t[k] - prev, (t[k] - prev) / 1965.0))
    prev = t[k]
print("rigid step: %.3f ms/step, %.0f env-steps/s (B=%d, %s, qp32=%s, mono=%s, inputs=%s)" % (ms, B / ms * 1e3, B, args.dtype, args.qp32, args.mono, args.inputs))
