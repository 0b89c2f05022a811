"""Phase timeline of one warp of step_kernel (clock64 marks, see B200SIM_PHASE_MARK):

    python scripts/phase_clocks.py [--batch 4096] [--dtype f32] [--lanes 8] [--generic]
"""
import argparse
import ctypes
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import jaxsim_b200.api as js  # noqa: E402
from jaxsim_b200 import _lib, models  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--model", default="icub_like")
ap.add_argument("--lanes", type=int, default=0)
ap.add_argument("--generic", action="store_true")
ap.add_argument("--no-caches", action="store_true")
ap.add_argument("--bulk-in", action="store_true")
ap.add_argument("--step-v1", action="store_true", help="first-generation specialised kernel (default: step2_kernel)")
ap.add_argument("--ring", type=int, default=1, help="number of independent state sets walked round-robin (>= 10 at batch 4096: inputs come from HBM)")
args = ap.parse_args()
dev = torch.device("cuda:0")
dtype = torch.float32 if args.dtype == "f32" else torch.float64
m = js.model.JaxSimModel.build_from_model_description(models.urdf(args.model), time_step=1e-3)
if args.lanes:
    m.set_tuning(lanes_per_env=args.lanes)
if args.generic or args.bulk_in or args.step_v1:
    m.set_options(generic_kernel=args.generic, bulk_in=args.bulk_in, step_v1=args.step_v1)
B, n = args.batch, m.dofs()
data = js.data.random_model_data(m, batch_size=B, seed=0, dtype=dtype, device=dev, velocity_representation=js.common.VelRepr.Inertial)
tau = 10 * torch.rand(B, n, dtype=dtype, device=dev)
out = js.model.step(m, data, joint_force_references=tau, update_caches=not args.no_caches)
lib = _lib.load()
h = m.device_model(dev).handle
cnt = (ctypes.c_ulonglong * 8)()
clk = (ctypes.c_ulonglong * 32)()
lib.b200sim_debug_counters.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
lib.b200sim_debug_phase_clocks.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
lib.b200sim_debug_counters(h, cnt)  # enable
NAMES = ["start", "model staged", "input burst issued", "base state ready", "inputs landed", "kinematics of input (cached unpack / jt+FK)",
         "contacts", "phase 3 link-parallel", "ABA pass 2", "base acceleration", "ABA pass 3", "Euler (base)", "base outputs stored",
         "joints + joint transforms + adjoint emit", "FK chain (new state)", "cache stores issued", "end (bulk stores drained)"]
if not (args.generic or args.step_v1):  # marks of step2_kernel (B200SIM_MARK2)
    NAMES = ["start", "model staged, previous launch complete", "input burst issued (2 bulk loads + joint state)", "base state ready", "inputs landed",
             "records from the cached rows", "contacts", "link-parallel (inertias, bias forces, actuation)", "ABA pass 2", "base acceleration",
             "ABA pass 3", "Euler (base) + base outputs", "joints: Euler update", "joint transforms, adjoints staged, bulk store",
             "FK walk (new state)", "link velocities + bulk stores of W_H_L / W_v", "end (bulk stores drained)"]
rows = []
for rep in range(5):
    for _ in range(3):
        js.model.step(m, data, joint_force_references=tau, out=out, update_caches=not args.no_caches)
    lib.b200sim_debug_phase_clocks(h, clk)
    rows.append(list(clk)[:17])
t = rows[-1]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    js.model.step(m, data, joint_force_references=tau, out=out, update_caches=not args.no_caches)
e1.record()
torch.cuda.synchronize()
print(f"{args.model} B={B} {args.dtype} lanes={args.lanes or 'auto'} generic={args.generic}: eager {e0.elapsed_time(e1) / 20 * 1e3:.2f} us/step")
# clock64 ticks at the SM clock
mhz = 1965.0
prev = t[0]
for k in range(1, 17):
    if t[k] == 0:
        continue
    d = t[k] - prev
    print(f"  {NAMES[k]:48s} {d:8d} cyc  {d / mhz:7.2f} us   (cum {(t[k] - t[0]) / mhz:6.2f} us)")
    prev = t[k]

import numpy as np  # noqa: E402

bt = (ctypes.c_ulonglong * 1536)()
lib.b200sim_debug_block_times.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
geo = m.launch_geometry(B, dtype, dev)
stats = []
ring = [js.data.random_model_data(m, batch_size=B, seed=100 + r, dtype=dtype, device=dev, velocity_representation=js.common.VelRepr.Inertial) for r in range(args.ring)]
outs = [js.model.step(m, d, joint_force_references=tau) for d in ring]
torch.cuda.synchronize()
for rep in range(max(5, 2 * args.ring)):
    r = rep % args.ring
    js.model.step(m, ring[r], joint_force_references=tau, out=outs[r], update_caches=not args.no_caches)
    lib.b200sim_debug_block_times(h, bt)
    a = np.array(list(bt)[:1024], dtype=np.float64).reshape(512, 2)[:min(512, geo["grid"])]
    smid = np.array(list(bt)[1024:], dtype=np.int64)[:min(512, geo["grid"])]
    t0 = a[:, 0].min()
    st, en = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3
    stats.append((st.max(), np.median(en - st), (en - st).min(), (en - st).max(), en.max(), np.percentile(en, 10), np.percentile(en, 90)))
dur = en - st
worst = np.argsort(-dur)[:6]
print("  slowest blocks of the last launch (block, smid, start, duration us):", [(int(b), int(smid[b]), round(float(st[b]), 2), round(float(dur[b]), 2)) for b in worst])
s_ = np.median(np.array(stats), axis=0)
print("  blocks (us, median of 5 launches): last block start %.2f | block duration median %.2f min %.2f max %.2f | last end %.2f (p10 %.2f p90 %.2f)" % tuple(s_))
