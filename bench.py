#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched simulation step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--batch B] [--dtype f32|f64]

Workload (config.workload): BASELINE.json configs[1] -- iCub-like 23-DoF floating-base
humanoid, soft contacts, fp32, batch 4096 PER GPU (weak scaling: every rank steps its own
shard of environments, no data-path collective; one NCCL all_gather of the state leaves is
performed after the timed region and reported separately as `readback_allgather_ms`).

One "step" = one call of the drop-in `step` over the whole batch.  Timing rules followed:
W >= 3 warm-ups; the K timed steps walk a RING of independent state sets whose total
footprint exceeds 2x the 126 MB L2, so every step's inputs come from HBM (config.l2);
CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
The K timed launches are replayed from one CUDA graph (launch-bound inner loop; `eager`
reports the same K steps launched one by one from Python).

Keys beyond the base contract: `roofline` (HBM, algorithmic bytes B_api of SURVEY.md 8d /
measured copy bandwidth), `cpu_baseline` (the NumPy oracle port timed on the host cores on
a bounded sample), `e2e` (same metric through the public API with pinned HOST buffers:
H2D of the state + torques, step, D2H of the new state, every step), `clocks`,
`gpu_launches`.
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "env-steps/sec"
UNIT = "env-steps/s"
WORKLOAD = "icub_like 23-DoF floating base, soft contacts (BASELINE.json configs[1])"
L2_BYTES = 126 * 1024 * 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--global-batch", type=int, default=0, help="total environments over all GPUs (overrides --batch: per-GPU batch = global / N); BASELINE configs[3] is --gpus 8 --global-batch 65536")
    ap.add_argument("--no-extras", action="store_true", help="skip the bounded extra legs of the default N=1 line (large-batch sweep, config 3, config 5, no-PDL)")
    ap.add_argument("--min-ms", type=float, default=50.0, help="the CUDA-graph of K steps is replayed until the timed region lasts at least this long")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--model", default="icub_like")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per env (0 = auto)")
    ap.add_argument("--epb", type=int, default=0, help="diagnostic: cap on the environments per thread block (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-caches", action="store_true", help="diagnostic: do not materialise the cached transforms (B_min traffic)")
    ap.add_argument("--skip-cache", default="", help="diagnostic: letters of caches NOT to write: X (joint transforms), H (link transforms), V (link velocities), B (base transform)")
    ap.add_argument("--out-ring", type=int, default=0, help="diagnostic: number of distinct output buffer sets (0 = same as the input ring)")
    ap.add_argument("--no-input-caches", action="store_true", help="recompute the kinematics of the input state instead of reading its cached link transforms/velocities")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    ap.add_argument("--no-tma", action="store_true", help="128-bit stores instead of TMA bulk stores for the joint adjoints")
    ap.add_argument("--no-pdl", action="store_true", help="diagnostic: ordinary launches instead of programmatic dependent launches")
    ap.add_argument("--bulk-in", action="store_true", help="diagnostic: cp.async.bulk (TMA) instead of per-link cp.async for the cached input kinematics")
    ap.add_argument("--step-v1", action="store_true", help="diagnostic: first-generation specialised step kernel instead of b200sim_step2")
    ap.add_argument("--no-bulk-in", action="store_true", help="diagnostic (step2): per-link cp.async instead of cp.async.bulk for the cached input kinematics")
    ap.add_argument("--generic-kernel", action="store_true", help="diagnostic: launch the generic step-kernel instance instead of the specialised one")
    ap.add_argument("--profile", action="store_true", help="cudaProfilerStart/Stop around the eager timed region (ncu --profile-from-start off)")
    ap.add_argument("--jvp", action="store_true", help="also time BASELINE config 5: forward-mode d(step)/d(joint q, link masses), fp64")
    ap.add_argument("--config3", action="store_true", help="also time BASELINE config 3: ErgoCub-like ~50-DoF, RIGID contacts, batch 16384 fp32")
    ap.add_argument("--c3-batch", type=int, default=16384)
    ap.add_argument("--rollout", type=int, default=0, help="also time step_n with this many fused steps per launch")
    ap.add_argument("--in-contact", action="store_true", help="also time the same batch with every environment touching the ground")
    ap.add_argument("--sweep", action="store_true", help="also time batch 16384 and 65536 on this GPU")
    return ap.parse_args()


def algorithmic_bytes_per_env(n, nL, nc, w, caches=True):
    """B_api / B_min of SURVEY.md 8d."""
    b = 2 * (2 * n + 13) + 6 * nc + n
    if caches:
        b += 16 + 58 * nL
    return w * b


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                if t0 - 0.2 <= t <= t1 + 0.2:
                    sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active") and t0 - 0.2 <= t <= t1 + 0.2:
                    reasons.add(name)
        if not sm:  # region shorter than the sampling period: use every sample we have
            for t, line in self.lines:
                try:
                    sm.append(float(line.split(",")[0]))
                except ValueError:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_oracle_throughput(model_name, batch, reps, threads):
    """The reference's CPU implementation of the path.  JAX/jaxsim cannot be installed
    offline, so this is the oracle PORT: the plain-C restatement
    (oracle/c/jaxsim_oracle.c, float64 = the reference's default dtype, dense 6x6 algebra as
    in the reference) on `threads` POSIX threads, each stepping a slice of the batch."""
    from oracle import c_oracle as CO
    from oracle import jaxsim_oracle as O
    from tests import helpers as H

    model = H.build_model(model_name)
    om = H.oracle_model(model)
    od = O.random_model_data(om, batch, seed=100)
    tau = 10 * np.random.default_rng(0).uniform(size=(batch, om.dofs()))
    CO.step(om, od, joint_force_references=tau, nthreads=threads)  # warm (also builds the library if needed)
    t0 = time.perf_counter()
    for _ in range(reps):
        out = CO.step(om, od, joint_force_references=tau, nthreads=threads)
    busy = time.perf_counter() - t0
    assert np.all(np.isfinite(out.joint_positions))
    return batch * reps / busy, busy


def run_reference(args):
    """`--impl reference`: see cpu_oracle_throughput.  One step = the full batch of the
    workload on all host cores; K steps, bounded to a few seconds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = max(1, min(os.cpu_count() or 1, 256))
    t = time.perf_counter()
    K = max(1, min(args.steps, 50))
    W = max(0, min(args.warmup, 3))
    if W:
        cpu_oracle_throughput(args.model, args.batch, W, threads)
    value, busy = cpu_oracle_throughput(args.model, args.batch, K, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * busy / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (random_model_data distribution, NumPy Philox)",
        "config": {"workload": WORKLOAD, "model": args.model, "batch_per_gpu": args.batch, "dt": 1e-3,
                   "note": "reference arm = plain-C oracle port of the jaxsim step (JAX is not installable offline); float64 like the reference's default"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"full {args.batch}-env batch x {K} steps on {threads} threads (C oracle port)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t,
    }
    emit(line)


def bind_to_gpu_numa_node(index: int) -> dict:
    """Pin this process to the CPU cores NVML reports as local to GPU `index` (its NUMA node) BEFORE any pinned host
    buffer is allocated: cudaHostAlloc places pages on the node of the calling thread, and a DMA from the far socket
    roughly halves the H2D / D2H rate of the e2e pipeline.  A no-op when the local cores are not in the cpuset."""
    info = {"bound": False}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        pick = sorted(local & allowed)
        info.update(gpu_local_cpus=len(local), allowed_cpus=len(allowed))
        if pick:
            os.sched_setaffinity(0, pick)
            info.update(bound=True, cpus=len(pick))
    except Exception as exc:  # NVML missing / not permitted: leave the affinity alone
        info["error"] = repr(exc)[:120]
    return info


# ----------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import jaxsim_b200.api as js
    from jaxsim_b200 import models

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    numa = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(dev)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    w = 4 if args.dtype == "f32" else 8

    model = js.model.JaxSimModel.build_from_model_description(models.urdf(args.model), time_step=1e-3)
    if args.lanes or args.epb:
        model.set_tuning(lanes_per_env=args.lanes, envs_per_block=args.epb)
    if args.no_tma or args.generic_kernel or args.bulk_in or args.no_pdl or args.step_v1 or args.no_bulk_in:
        model.set_options(tma_store=not args.no_tma, generic_kernel=args.generic_kernel, bulk_in=args.bulk_in, pdl=not args.no_pdl,
                          step_v1=args.step_v1, no_bulk_in=args.no_bulk_in)
    n, nL, nc = model.dofs(), model.number_of_links(), model.number_of_collidable_points()
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit("--global-batch must be a multiple of the number of GPUs")
        args.batch = args.global_batch // world
    B = args.batch
    extras = (world == 1) and not args.no_extras and args.model == "icub_like" and args.dtype == "f32" and not args.no_caches
    bytes_env = algorithmic_bytes_per_env(n, nL, nc, w, caches=not args.no_caches)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def lowered_to_contact(d, seed):
        """The same random state with the base lowered so that the lowest collidable point penetrates the
        ground by 0-5 mm (the second input distribution of BASELINE.md: every environment in contact)."""
        gen = torch.Generator(device=dev).manual_seed(seed)
        z = js.contact.collidable_point_positions(model, d)[..., 2].min(dim=-1).values
        p = d._base_position.clone()
        p[:, 2] -= z + 0.005 * torch.rand(p.shape[0], dtype=dtype, device=dev, generator=gen)
        return js.data.JaxSimModelData.build(
            model, base_position=p, base_quaternion=d._base_quaternion, joint_positions=d._joint_positions,
            joint_velocities=0.1 * d._joint_velocities, base_linear_velocity=0.1 * d._base_linear_velocity,
            base_angular_velocity=0.1 * d._base_angular_velocity, velocity_representation=js.common.VelRepr.Inertial,
            batch_size=p.shape[0], dtype=dtype, device=dev)

    def time_steps(Bq, K, W, use_graph=True, in_contact=False, model=model):
        """K device-resident steps over a ring of state sets larger than L2.  Returns
        (ms of the K steps launched eagerly, ms of the same K steps replayed from a CUDA
        graph or None, ring, last output).  The graph of K steps is replayed back to back until the
        timed region lasts at least --min-ms (the figure is the mean over the replays)."""
        ring = max(2, int(np.ceil(2 * L2_BYTES / (Bq * bytes_env))))
        ring = min(ring, 64)
        datas = [js.data.random_model_data(model, batch_size=Bq, seed=1000 * rank + r, dtype=dtype, device=dev,
                                           velocity_representation=js.common.VelRepr.Inertial) for r in range(ring)]
        if in_contact:
            datas = [lowered_to_contact(d, 77 + r) for r, d in enumerate(datas)]
        taus = [10 * torch.rand(Bq, n, dtype=dtype, device=dev) for _ in range(ring)]
        # preallocated outputs (`out=`): the step then performs no allocation and is capturable
        outs = [js.model.step(model, datas[r], joint_force_references=taus[r], update_caches=not args.no_caches) for r in range(ring)]

        for o in outs:
            if "X" in args.skip_cache: o._joint_transforms = None
            if "H" in args.skip_cache: o._link_transforms = None
            if "V" in args.skip_cache: o._link_velocities = None
            if "B" in args.skip_cache: o._base_transform = None
        oring = args.out_ring if args.out_ring > 0 else ring

        def run(count):
            o = None
            for i in range(count):
                o = js.model.step(model, datas[i % ring], joint_force_references=taus[i % ring], out=outs[i % oring],
                                  update_caches=not args.no_caches, use_input_caches=not args.no_input_caches)
            return o

        run(W)
        barrier()
        if args.profile:
            torch.cuda.profiler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = run(K)
        e1.record()
        barrier()
        if args.profile:
            torch.cuda.profiler.stop()
        ms_eager = e0.elapsed_time(e1)
        ms_graph = None
        if use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run(K)
            g.replay()
            barrier()
            e0.record()
            g.replay()
            e1.record()
            barrier()
            once = e0.elapsed_time(e1)
            reps = int(min(2000, max(1, np.ceil(args.min_ms / max(once, 1e-3)))))
            if world > 1:  # the same number of replays on every rank
                tr_ = torch.tensor([reps], dtype=torch.int64, device=dev)
                dist.all_reduce(tr_, op=dist.ReduceOp.MAX)
                reps = int(tr_.item())
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            barrier()
            ms_graph = e0.elapsed_time(e1) / reps
            time_steps.replays = reps
        return ms_eager, ms_graph, ring, out

    time_steps.replays = 1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t0 = time.time()
    ms_eager, ms_graph, ring, out = time_steps(B, args.steps, max(3, args.warmup), use_graph=not args.no_graph)
    ms = ms_graph if ms_graph is not None else ms_eager
    main_replays = time_steps.replays
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None

    t = torch.tensor([ms, ms_eager], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, ms_eager_max = float(t[0].item()), float(t[1].item())
    value = B * world * args.steps / (ms_max * 1e-3)

    # ---- state readback across ranks (config 4: NCCL all_gather only at readback)
    gather_ms = None
    if world > 1:
        from jaxsim_b200.distributed import pack_state

        leaves = pack_state(out)
        full = torch.empty(world * B, leaves.shape[1], dtype=dtype, device=dev)
        dist.all_gather_into_tensor(full, leaves)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(full, leaves)
        g1.record()
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather_ms = float(tg.item())

    # ---- e2e: public API with pinned HOST buffers; every step = H2D of that step's inputs
    # (state, contact state, joint force references), js.model.step, D2H of the new state.
    # The leaves of a step live in ONE flat buffer per direction (one memcpy each way) and
    # the three stages are double-buffered on three streams, as a host-driven rollout would.
    Ke = max(5, min(args.steps, 50))
    r4 = lambda k: (k + 3) & ~3  # noqa: E731
    in_blocks = [("s", (B, n)), ("sd", (B, n)), ("q", (B, 4)), ("vl", (B, 3)), ("om", (B, 3)), ("p", (B, 3)),
                 ("m", (B, nc, 3)), ("tau", (B, n))]
    out_blocks = in_blocks[:-1]

    def carve(flat, blocks):
        views, o = {}, 0
        for name, shape in blocks:
            k = int(np.prod(shape))
            views[name] = flat[o:o + k].view(shape)
            o += r4(k)
        return views

    n_in = sum(r4(int(np.prod(s))) for _, s in in_blocks)
    n_out = sum(r4(int(np.prod(s))) for _, s in out_blocks)
    src = js.data.random_model_data(model, batch_size=B, seed=7 + rank, dtype=dtype, device=dev,
                                    velocity_representation=js.common.VelRepr.Inertial)
    h_in = torch.empty(n_in, dtype=dtype).pin_memory()
    hv = carve(h_in, in_blocks)
    for k, leaf in (("s", "_joint_positions"), ("sd", "_joint_velocities"), ("q", "_base_quaternion"),
                    ("vl", "_base_linear_velocity"), ("om", "_base_angular_velocity"), ("p", "_base_position")):
        hv[k].copy_(getattr(src, leaf).cpu())
    hv["m"].zero_()
    hv["tau"].copy_(10 * torch.rand(B, n, dtype=dtype))
    NB = max(2, int(os.environ.get("B200SIM_E2E_BUFFERS", "2")))  # pipeline depth (buffers per stage)
    h_out = [torch.empty(n_out, dtype=dtype).pin_memory() for _ in range(NB)]
    d_in = [torch.empty(n_in, dtype=dtype, device=dev) for _ in range(NB)]
    d_out = [torch.empty(n_out, dtype=dtype, device=dev) for _ in range(NB)]

    def mk(views, with_tau):
        d = js.data.JaxSimModelData(
            velocity_representation=js.common.VelRepr.Inertial, _joint_positions=views["s"], _joint_velocities=views["sd"],
            _base_quaternion=views["q"], _base_linear_velocity=views["vl"], _base_angular_velocity=views["om"],
            _base_position=views["p"], contact_state={"tangential_deformation": views["m"]})
        return (d, views["tau"]) if with_tau else d

    din = [mk(carve(d_in[j], in_blocks), True) for j in range(NB)]
    dout = [mk(carve(d_out[j], out_blocks), False) for j in range(NB)]
    for d in dout:  # the step does the same device work as the `value` arm: caches are written (not copied back)
        d._base_transform = torch.empty(B, 4, 4, dtype=dtype, device=dev)
        d._joint_transforms = torch.empty(B, nL, 6, 6, dtype=dtype, device=dev)
        d._link_transforms = torch.empty(B, nL, 4, 4, dtype=dtype, device=dev)
        d._link_velocities = torch.empty(B, nL, 6, dtype=dtype, device=dev)
    h2d = n_in * h_in.element_size()
    d2h = n_out * h_in.element_size()
    s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(NB)]
    ev_cmp = [torch.cuda.Event() for _ in range(NB)]
    ev_out = [torch.cuda.Event() for _ in range(NB)]

    def e2e_run(count, fresh=False, e2e_nostep=False):
        """`count` pipelined steps over three streams.  `fresh`: no event has been recorded yet
        (first use, or inside a graph capture where only captured events may be waited on)."""
        seen = set() if fresh else {(k, j) for k in ("cmp", "out") for j in range(NB)}
        for i in range(count):
            j = i % NB
            with torch.cuda.stream(s_in):
                if ("cmp", j) in seen:
                    s_in.wait_event(ev_cmp[j])       # the step that last read d_in[j] is done
                d_in[j].copy_(h_in, non_blocking=True)
                ev_in[j].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[j])
                if ("out", j) in seen:
                    s_cmp.wait_event(ev_out[j])      # the D2H that last read d_out[j] is done
                if not e2e_nostep:
                    js.model.step(model, din[j][0], joint_force_references=din[j][1], out=dout[j])
                ev_cmp[j].record(s_cmp)
                seen.add(("cmp", j))
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[j])
                h_out[j].copy_(d_out[j], non_blocking=True)
                ev_out[j].record(s_out)
                seen.add(("out", j))

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_cmp)
        fn()
        e1.record(s_cmp)
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())

    # (1) launched call by call from Python (host dispatch of 2 copies + 1 step per iteration included)
    e2e_run(2 * NB, fresh=True)

    def eager():
        e2e_run(Ke)
        for j in range(NB):
            s_cmp.wait_event(ev_out[j])

    e2e_eager_value = B * world * Ke / (timed(eager) * 1e-3)

    def copies_only():
        e2e_run(Ke, e2e_nostep=True)
        for j in range(NB):
            s_cmp.wait_event(ev_out[j])

    copies_only()
    e2e_copy_only_value = B * world * Ke / (timed(copies_only) * 1e-3)

    # (2) the same Ke-step pipeline (same API calls, same pinned buffers, same three streams) captured
    #     once into a CUDA graph and replayed: what a user does to take Python out of the loop
    e2e_value, e2e_mode = e2e_eager_value, "eager"
    if not args.no_graph:
        try:
            torch.cuda.synchronize()
            eg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(eg, stream=s_cmp):
                s_in.wait_stream(s_cmp)
                s_out.wait_stream(s_cmp)
                e2e_run(Ke, fresh=True)
                s_cmp.wait_stream(s_in)
                s_cmp.wait_stream(s_out)
            with torch.cuda.stream(s_cmp):
                eg.replay()
                torch.cuda.synchronize()
                t_graph = timed(eg.replay)
            e2e_graph_value = B * world * Ke / (t_graph * 1e-3)
            if e2e_graph_value >= e2e_eager_value:
                e2e_value, e2e_mode = e2e_graph_value, "cuda_graph"
            else:  # the driver sometimes serialises the three captured streams: the eager pipeline is the better user path then
                e2e_value, e2e_mode = e2e_eager_value, "eager (faster than the captured pipeline: %.3e)" % e2e_graph_value
        except Exception as exc:  # a failed capture must not cost the run its result line: keep the eager figure
            print(f"[bench] e2e graph capture failed ({exc!r}); reporting the eagerly dispatched pipeline", file=sys.stderr)
            torch.cuda.synchronize()
            e2e_value, e2e_mode = e2e_eager_value, "eager (graph capture failed)"

    # ---- optional batch sweep on this GPU (metric is quoted "batch 4096 -> 65536")
    sweep = None
    peak_gbs, _ = measured_peak_gbs()
    if (args.sweep or extras) and world == 1:
        sweep = []
        for Bq in (16384, 65536):
            Kq = 20
            mse, msg, rq, _ = time_steps(Bq, Kq, 3, use_graph=not args.no_graph)
            msq = msg if msg is not None else mse
            ach = Bq * bytes_env / (msq / Kq * 1e-3) / 1e9
            sweep.append({"batch": Bq, "value": Bq * Kq / (msq * 1e-3), "ms_per_step": msq / Kq, "ring": rq, "steps": Kq,
                          "replays": time_steps.replays, "launch": model.launch_geometry(Bq, dtype, dev),
                          "roofline": {"bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs}})

    # the same default workload launched WITHOUT programmatic dependent launch (closes the loop on the
    # launch-to-launch figure being shorter than the kernel duration ncu reports)
    no_pdl = None
    if extras and not args.no_pdl and not args.no_graph:
        m2 = js.model.JaxSimModel.build_from_model_description(models.urdf(args.model), time_step=1e-3)
        m2.set_options(pdl=False, step_v1=args.step_v1, no_bulk_in=args.no_bulk_in)
        mse, msg, _, _ = time_steps(B, min(args.steps, 50), 3, model=m2)
        no_pdl = {"ms_per_step": msg / min(args.steps, 50), "eager_ms_per_step": mse / min(args.steps, 50),
                  "note": "same K-step CUDA graph with ordinary launches (B200SIM_OPT_NO_PDL): the difference to ms_per_step is the launch set-up hidden by programmatic dependent launch"}

    contact_leg = None
    if (args.sweep or args.in_contact or extras) and world == 1 and nc > 0:
        Kc = max(20, args.steps // 4)
        mse, msg, rq, _ = time_steps(B, Kc, 3, use_graph=not args.no_graph, in_contact=True)
        msc = msg if msg is not None else mse
        contact_leg = {"batch": B, "value": B * Kc / (msc * 1e-3), "ms_per_step": msc / Kc,
                       "inputs": "every environment touching the ground (lowest point 0-5 mm below it); the default line uses the "
                                 "reference benchmark's random_model_data, where most environments are airborne"}

    rollout = None
    if args.rollout > 0 or extras:
        Tn = args.rollout if args.rollout > 0 else 32
        d0 = js.data.random_model_data(model, batch_size=B, seed=55 + rank, dtype=dtype, device=dev,
                                       velocity_representation=js.common.VelRepr.Inertial)
        tau_T = 10 * torch.rand(Tn, B, n, dtype=dtype, device=dev)
        o = js.model.step_n(model, d0, Tn, joint_force_references=tau_T, update_caches=False)
        for _ in range(3):
            js.model.step_n(model, d0, Tn, joint_force_references=tau_T, update_caches=False, out=o)
        barrier()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            js.model.step_n(model, d0, Tn, joint_force_references=tau_T, update_caches=False, out=o)
        e1.record()
        barrier()
        tr = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        rollout = {"steps_per_launch": Tn, "value": B * world * Tn * reps / (float(tr.item()) * 1e-3), "unit": UNIT,
                   "ms_per_step": float(tr.item()) / (reps * Tn),
                   "note": "step_n (SURVEY.md 8f-1): the user loop `for _ in range(T): data = step(model, data, tau[t])` as ONE launch, "
                           "state kept on chip, per-step HBM traffic = that step's joint force references; no caches written (B_min regime)"}

    config3 = None
    relaxed3 = None

    def time_contact_model(kind):
        from jaxsim_b200.rbda.contacts import RelaxedRigidContacts, RelaxedRigidContactsParams, RigidContacts, RigidContactsParams

        if kind == "rigid":
            cm3, cp3 = RigidContacts.build(), RigidContactsParams.build(K=1e4, D=20.0)
        else:  # the contact model of the reference's own step benchmark (tests/test_benchmark.py:143-152)
            cm3, cp3 = RelaxedRigidContacts.build(), RelaxedRigidContactsParams.build(mu=0.5)
        m3 = js.model.JaxSimModel.build_from_model_description(
            models.urdf("ergocub_like"), time_step=1e-3, contact_model=cm3, contact_params=cp3)
        B3 = args.c3_batch
        n3, nL3, nc3 = m3.dofs(), m3.number_of_links(), m3.number_of_collidable_points()
        mono3 = bool(os.environ.get("B200SIM_RIGID_MONO"))  # A/B switch: contact QP inside the rigid kernel (round-1 path)
        if mono3:
            m3.set_options(rigid_mono=True)

        def standing(seed):
            """level base, near-zero joints, soles 2-5 mm into the ground: several points active"""
            gen = torch.Generator(device=dev).manual_seed(seed)
            u = lambda *sh: 2 * torch.rand(*sh, dtype=dtype, device=dev, generator=gen) - 1  # noqa: E731
            rpy = 1e-3 * u(B3, 3)
            q = torch.cat([torch.ones(B3, 1, dtype=dtype, device=dev), 0.5 * rpy], dim=-1)
            q = q / q.norm(dim=-1, keepdim=True)
            p = torch.cat([u(B3, 2), torch.ones(B3, 1, dtype=dtype, device=dev)], dim=-1)
            kw = dict(base_quaternion=q, joint_positions=1e-3 * u(B3, n3), joint_velocities=0.1 * u(B3, n3),
                      base_linear_velocity=0.1 * u(B3, 3), base_angular_velocity=0.1 * u(B3, 3),
                      velocity_representation=js.common.VelRepr.Inertial, batch_size=B3, dtype=dtype, device=dev)
            d0 = js.data.JaxSimModelData.build(m3, base_position=p, **kw)
            cp = m3.kin_dyn_parameters.contact_parameters
            body = torch.as_tensor(np.array(cp.body), device=dev)
            Lp = torch.as_tensor(np.asarray(cp.point), dtype=dtype, device=dev)
            H = d0.link_transforms[:, body]
            z = (H[..., 2, 0:3] * Lp).sum(-1) + H[..., 2, 3]
            drop = z.min(dim=1).values + 0.002 + 0.003 * torch.rand(B3, dtype=dtype, device=dev, generator=gen)
            p = p.clone()
            p[:, 2] -= drop
            d1 = js.data.JaxSimModelData.build(m3, base_position=p, **kw)
            H = d1.link_transforms[:, body]
            act = ((H[..., 2, 0:3] * Lp).sum(-1) + H[..., 2, 3] < 0).sum(dim=1).float().mean().item()
            return d1, act

        config3 = {"config": "%s: ergocub_like (%d DoF, %d links, %d collidable points), %s, batch %d %s"
                             % ("BASELINE configs[2]" if kind == "rigid" else "reference step benchmark (tests/test_benchmark.py:143-152)",
                                n3, nL3, nc3, type(cm3).__name__, B3, args.dtype), "unit": "env-steps/s"}
        if kind == "rigid":
            config3["level1"] = ("monolithic rigid kernel (B200SIM_OPT_RIGID_MONO)" if mono3 else
                                 "split: assemble (rigid kernel) / solve (rigid_qp_kernel, QPs of <= 8 and 9-12 active points) / resume (rigid kernel)")
        ring3 = 4
        for label in ("random", "standing"):
            if label == "random":
                ds = [js.data.random_model_data(m3, batch_size=B3, seed=50 + r + 1000 * rank, dtype=dtype, device=dev,
                                                velocity_representation=js.common.VelRepr.Inertial) for r in range(ring3)]
                act = None
            else:
                pairs = [standing(70 + r + 1000 * rank) for r in range(ring3)]
                ds, act = [p_[0] for p_ in pairs], float(np.mean([p_[1] for p_ in pairs]))
            ts = [10 * torch.rand(B3, n3, dtype=dtype, device=dev) for _ in range(ring3)]
            os_ = [js.model.step(m3, ds[r], joint_force_references=ts[r]) for r in range(ring3)]
            K3 = max(4, min(args.steps, 20))
            for i in range(3):
                js.model.step(m3, ds[i % ring3], joint_force_references=ts[i % ring3], out=os_[i % ring3])
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K3):
                js.model.step(m3, ds[i % ring3], joint_force_references=ts[i % ring3], out=os_[i % ring3])
            e1.record()
            barrier()
            t3 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            ms3 = float(t3.item()) / K3
            b3 = algorithmic_bytes_per_env(n3, nL3, 0, w)  # rigid contacts carry no contact state
            ach3 = B3 * b3 / (ms3 * 1e-3) / 1e9
            config3[label] = {"ms_per_step": ms3, "value": B3 * world / (ms3 * 1e-3), "steps": K3,
                              "mean_active_points": act,
                              "roofline": {"bound": "hbm", "achieved": ach3, "peak": peak_gbs, "unit": "GB/s", "frac": ach3 / peak_gbs,
                                           "bytes_per_env_step": b3,
                                           "note": "B_api of SURVEY.md 8d for this model; the rigid step is bound by the warp-serial contact QP, not by HBM"},
                              "inputs": "random_model_data (base 0.5-1 m above ground: mostly no contact)" if label == "random"
                              else "standing: level base, soles 2-5 mm into the ground"}
        if kind == "rigid" and rank == 0 and not args.no_cpu_baseline:
            config3["cpu_baseline"] = cpu_rigid_oracle_throughput(standing)
        return config3

    def cpu_rigid_oracle_throughput(standing, sample=128):
        """CPU arm of config 3: the NumPy restatement of the reference's rigid-contact step (oracle/rigid_oracle.py,
        float64, vectorised over the batch; NumPy/BLAS threads as configured) on a bounded sample of the SAME
        `standing` inputs."""
        from oracle import jaxsim_oracle as O3
        from oracle import rigid_oracle as R3
        from tests import helpers as H3
        from jaxsim_b200.rbda.contacts import RigidContacts, RigidContactsParams

        m3 = H3.build_model("ergocub_like", contact_model=RigidContacts.build(), contact_params=RigidContactsParams.build(K=1e4, D=20.0))
        om3 = H3.oracle_model(m3)
        d1, _ = standing(70)
        c = lambda t: t[:sample].detach().cpu().double().numpy()  # noqa: E731
        od3 = O3.data_replace(om3, c(d1._joint_positions), c(d1._joint_velocities), c(d1._base_quaternion),
                              c(d1._base_linear_velocity), c(d1._base_angular_velocity), c(d1._base_position))
        tau3 = 10 * np.random.default_rng(0).uniform(size=(sample, om3.dofs()))
        t0_ = time.perf_counter()
        R3.step(om3, od3, joint_force_references=tau3)
        busy = time.perf_counter() - t0_
        return {"value": sample / busy, "unit": UNIT, "cores": int(os.cpu_count() or 1), "kind": "port",
                "sample": f"{sample} 'standing' environments x 1 step, NumPy restatement of the reference's rigid-contact step (float64)"}

    def guarded(label, fn):
        """An extra leg must never cost the run its headline line: a failure is recorded in the leg's place."""
        try:
            return fn()
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] leg {label} failed: {exc!r}", file=sys.stderr)
            try:
                torch.cuda.synchronize()
            except Exception:  # noqa: BLE001
                pass
            return {"error": repr(exc)}

    if args.config3 or extras:
        config3 = guarded("config3_rigid", lambda: time_contact_model("rigid"))
    if args.config3:
        relaxed3 = guarded("relaxed_rigid", lambda: time_contact_model("relaxed"))

    jvp = None

    def time_jvp():
        d64 = js.data.random_model_data(model, batch_size=B, seed=77 + rank, dtype=torch.float64, device=dev,
                                        velocity_representation=js.common.VelRepr.Inertial)
        tq = torch.randn(B, n, dtype=torch.float64, device=dev)
        tm = torch.rand(nL, dtype=torch.float64)
        for _ in range(3):
            js.model.step_jvp(model, d64, {"joint_positions": tq, "link_masses": tm})
        barrier()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            js.model.step_jvp(model, d64, {"joint_positions": tq, "link_masses": tm})
        e1.record()
        barrier()
        ms_j = e0.elapsed_time(e1) / reps
        from jaxsim_b200.api import autodiff

        autodiff.step_jacobian(model, d64, ("joint_positions", "link_masses"))
        barrier()
        e0.record()
        for _ in range(3):
            autodiff.step_jacobian(model, d64, ("joint_positions", "link_masses"))
        e1.record()
        barrier()
        ms_jac = e0.elapsed_time(e1) / 3
        n_out = 2 * n + 13 + 3 * nc
        ct = torch.randn(B, n_out, dtype=torch.float64, device=dev)
        autodiff.step_vjp(model, d64, ct)
        barrier()
        e0.record()
        for _ in range(3):
            autodiff.step_vjp(model, d64, ct)
        e1.record()
        barrier()
        ms_vjp = e0.elapsed_time(e1) / 3
        return {"config": "BASELINE configs[4]: icub_like fp64, d(step)/d(joint q, link masses), batch %d" % B,
               "ms_per_jvp": ms_j, "env_jvps_per_s": B / (ms_j * 1e-3),
               "ms_full_jacobian": ms_jac, "jacobian_columns": n + nL, "ms_vjp": ms_vjp,
               "note": "ms_per_jvp: one tangent direction, value + tangent of every output leaf incl. caches.  ms_full_jacobian: "
                       "MEASURED autodiff.step_jacobian w.r.t. (joint positions, link masses) -> (B, n_out, n + nL): the n joint "
                       "directions and the nL link-mass directions as launches over replicas of the batch (b200sim_step_jvp_ex), no caches.  ms_vjp: b200sim_step_vjp, the "
                       "gradient of <cotangent, step> w.r.t. (joint positions, link masses): the same columns contracted on the device"}

    if args.jvp or extras:
        jvp = guarded("config5_jvp", time_jvp)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    ms_step = ms_max / args.steps
    achieved = B * bytes_env / (ms_step * 1e-3) / 1e9
    geo = model.launch_geometry(B, dtype, dev)
    # DRAM traffic and counted arithmetic of this configuration come from the committed ncu counter capture of the same
    # command line (scripts/step_counters.py -> profiles/r02_step_counters.json): steady-state DRAM bytes per launch
    # (reads + the drained writes, mean over consecutive launches on a ring of inputs larger than L2) and
    # 2*FFMA + FADD + FMUL thread instructions per environment step.  None when no capture exists for the configuration.
    traffic, compute = None, None
    tp = ROOT / "profiles" / "r02_step_counters.json"
    if tp.exists() and not args.no_caches:
        try:
            rec = json.loads(tp.read_text()).get(f"{args.model}_{args.dtype}_B{B}")
        except Exception:
            rec = None
        if rec:
            traffic = rec.get("dram_bytes_per_launch")
            fpe = rec.get("flops_per_env_step")
            if fpe and args.dtype == "f32":
                peak_fp32 = 148 * 128 * 2 * 1.965e9 / 1e12  # CUDA-core FMA peak at the maximum SM clock (nominal, TFLOP/s)
                compute = {"flops_per_env_step": fpe, "achieved_tflops": fpe * B / (ms_step * 1e-3) / 1e12, "peak_tflops": peak_fp32,
                           "frac": fpe * B / (ms_step * 1e-3) / 1e12 / peak_fp32,
                           "unit": "TFLOP/s fp32 (CUDA cores; tensor cores do not apply)", "source": rec.get("source")}

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # contract: the CPU baseline is timed at N = 1 only
        threads = max(1, min(os.cpu_count() or 1, 256))
        v, busy = cpu_oracle_throughput(args.model, B, 10, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"full {B}-env batch x 10 steps on {threads} threads (plain-C oracle port, float64; JAX unavailable offline)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic (random_model_data distribution, seeded torch generator; random-init state)",
        "config": {"workload": WORKLOAD, "model": args.model, "dofs": n, "links": nL, "collidable_points": nc,
                   "batch_per_gpu": B, "global_batch": B * world, "dt": 1e-3, "contact_model": "soft",
                   "integrator": "semi_implicit_euler", "parallelism": f"env-parallel x{world} (no data-path collective)",
                   "l2": f"inputs larger than L2: ring of {ring} independent state sets ({ring * B * bytes_env / 2**20:.0f} MiB)",
                   "launch": geo, "caches_written": not args.no_caches, "input_caches_read": not args.no_input_caches, "cuda_graph": ms_graph is not None,
                   "joint_adjoint_store": "128-bit STG" if args.no_tma else "TMA cp.async.bulk"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "bytes_per_env_step": bytes_env, "peak_source": peak_src,
                     "note": "algorithmic bytes = B_api (SURVEY.md 8d): read state+contact state+tau, write state+contact state+all caches; the kernel additionally READS the input data's cached link transforms/velocities (88 B/link, +2112 B/env) instead of recomputing them, like the reference's contact code",
                     "bytes_moved_per_env_step": bytes_env + (0 if (args.no_input_caches or args.no_caches) else w * 22 * nL)},
        "compute": compute,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                "launch": e2e_mode, "eager_value": e2e_eager_value, "copy_only_value": e2e_copy_only_value, "numa": numa,
                "note": "public API js.model.step with pinned host buffers, every step: one H2D (state + contact state + tau), step, one D2H (new state + contact state); pipelined over 3 streams (%d buffers per stage); the caches are written on the device like in `value` but not copied back.  `value`: the Ke-step pipeline captured into one CUDA graph and replayed (launch=cuda_graph); `eager_value`: the same calls dispatched from Python one by one" % NB},
        "gpu_launches": args.steps * main_replays, "replays": main_replays,
        "eager": {"value": B * world * args.steps / (ms_eager_max * 1e-3), "ms_per_step": ms_eager_max / args.steps,
                  "note": "same K steps launched one by one from Python (host launch latency included)"},
        "clocks": clocks,
    }
    if gather_ms is not None:
        line["readback_allgather_ms"] = gather_ms
    if sweep is not None:
        line["large_batch"] = sweep
    if no_pdl is not None:
        line["no_pdl"] = no_pdl
    if contact_leg is not None:
        line["in_contact"] = contact_leg
    if rollout is not None:
        line["rollout"] = rollout
    if jvp is not None:
        line["config5_jvp"] = jvp
    if config3 is not None:
        line["config3_rigid"] = config3
    if relaxed3 is not None:
        line["relaxed_rigid"] = relaxed3
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    args = parse()
    # Libraries may write to fd 1 (NCCL prints its version banner there when NCCL_DEBUG is set):
    # keep the original stdout for the result line only and send everything else to stderr.
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
