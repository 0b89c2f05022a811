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
args = ap.parse_args()
dev = torch.device("cuda:0")
dtype = torch.float32 if args.dtype == "f32" else torch.float64
m = js.model.JaxSimModel.build_from_model_description(models.urdf(args.model), time_step=1e-3)
if args.lanes:
    m.set_tuning(lanes_per_env=args.lanes)
if args.generic:
    m.set_options(generic_kernel=True)
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
