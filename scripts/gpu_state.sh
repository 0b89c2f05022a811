#!/bin/bash
# state-of-the-repo GPU round: parity suite, default bench, reference arm, config 3, rollout/sweep, ncu launch list + full capture
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
nvidia-smi -L | head -2; nproc
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== bench (default)"
timeout 600 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_default.json | cut -c1-3000
tail -5 gpurun_out/bench_err.log
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_reference.json | cut -c1-600
echo "== bench config3 + rollout + sweep + jvp"
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --config3 --rollout 100 --sweep --jvp 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_extras.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
for k in ('config3_rigid','rollout','sweep','jvp'): print(k, json.dumps(d.get(k))[:1500])"
echo "== rigid profile driver"
timeout 300 python scripts/rigid_profile.py --batch 16384 --steps 5 2>&1 | tail -3
echo "== ncu launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph --profile > gpurun_out/ncu_launch_bench.log 2>&1
grep -c step_kernel gpurun_out/launches.csv; tail -2 gpurun_out/launches.csv | cut -c1-300
echo "== ncu full"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:step_kernel -c 2 -f -o gpurun_out/prof_step \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --profile > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out | tail -20
