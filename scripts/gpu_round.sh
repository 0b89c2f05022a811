#!/bin/bash
# One gpurun call = tests + bench + tuning sweep + ncu captures (box acquisition is charged
# per call, so everything that needs a GPU is batched here).  Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== nvidia-smi"; nvidia-smi -L; nproc
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench (default)"
timeout 600 python bench.py --steps 200 --warmup 10 2>gpurun_out/bench_err.log | tee gpurun_out/bench_default.json
tail -3 gpurun_out/bench_err.log
echo "== lanes sweep (CUDA-graph timing)"
for G in 4 8; do
  for B in 4096 8192 16384 65536; do
    timeout 300 python bench.py --steps 100 --warmup 5 --lanes $G --batch $B --no-cpu-baseline 2>>gpurun_out/sweep_err.log \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('G=$G B=$B', 'us/step graph=%.2f eager=%.2f'%(1e3*d['ms_per_step'],1e3*d['eager']['ms_per_step']), 'Menv/s=%.1f'%(d['value']/1e6), 'hbm_frac=%.3f'%d['roofline']['frac'], 'e2e Menv/s=%.2f'%(d['e2e']['value']/1e6), d['config']['launch'])" \
      | tee -a gpurun_out/lanes_sweep.log
  done
done
for EXTRA in "--no-tma" "--no-caches"; do
    timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline $EXTRA 2>>gpurun_out/sweep_err.log \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B=4096 $EXTRA', 'us/step graph=%.2f eager=%.2f'%(1e3*d['ms_per_step'],1e3*d['eager']['ms_per_step']), 'Menv/s=%.1f'%(d['value']/1e6), 'hbm_frac=%.3f'%d['roofline']['frac'])" \
      | tee -a gpurun_out/lanes_sweep.log
done
echo "== fp64 + rollout"
timeout 300 python bench.py --steps 100 --warmup 5 --dtype f64 --no-cpu-baseline --rollout 50 2>>gpurun_out/sweep_err.log | tee gpurun_out/bench_f64.json | cut -c1-300
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --rollout 100 --sweep 2>>gpurun_out/sweep_err.log | tee gpurun_out/bench_rollout_sweep.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['rollout']); print(d['sweep'])"
echo "== ncu launch list (timed region only)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph --profile > gpurun_out/ncu_launch_bench.log 2>&1
grep -c step_kernel gpurun_out/launches.csv; tail -2 gpurun_out/launches.csv | cut -c1-300
echo "== ncu full (timed region only)"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:step_kernel -c 2 -f -o gpurun_out/prof_step \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --profile > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out | tail -20
echo "== compute-sanitizer (memcheck + racecheck + synccheck) on the smoke step"
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke" gpurun_out/sanitizer_$tool.log | tail -4
done
