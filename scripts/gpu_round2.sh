#!/bin/bash
# consolidated GPU round: parity suite, default bench + reference arm, extras, ncu launch list, ncu full (cold + warm caches),
# phase clocks, sanitizers with and without the bulk-load path
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
nvidia-smi -L | head -1; nproc
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | cut -c1-2000 | tee gpurun_out/pytest_gpu.log
echo "== bench (default)"
timeout 600 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_default.json | cut -c1-3500
tail -5 gpurun_out/bench_err.log
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_reference.json | cut -c1-400
echo "== bench extras"
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --config3 --rollout 100 --sweep 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_extras.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
for k in ('config3_rigid','relaxed_rigid','rollout','sweep','in_contact'): print(k, json.dumps(d.get(k))[:1500])"
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --dtype f64 --jvp 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_f64.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('f64 us/step', 1e3*d['ms_per_step'], 'frac', d['roofline']['frac']); print('jvp', json.dumps(d.get('config5_jvp'))[:800])"
echo "== phase clocks"
python scripts/phase_clocks.py 2>&1 | tee gpurun_out/phase_clocks.log
python scripts/phase_clocks.py --ring 12 2>&1 | tail -2 | tee -a gpurun_out/phase_clocks.log
echo "== ncu launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph --profile > gpurun_out/ncu_launch_bench.log 2>&1
grep -c step_kernel gpurun_out/launches.csv; tail -1 gpurun_out/launches.csv | cut -c1-300
echo "== ncu full (cold caches, default) and (warm caches)"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:step_kernel -c 2 -f -o gpurun_out/prof_step \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --profile > gpurun_out/ncu_full_bench.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --cache-control none --import-source on -k regex:step_kernel -c 2 -f -o gpurun_out/prof_step_warm \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --profile > gpurun_out/ncu_full_warm_bench.log 2>&1
ls -la gpurun_out | tail -8
echo "== sanitizers (smoke)"
for tool in racecheck memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke" gpurun_out/sanitizer_$tool.log | head -4
done
B200SIM_SMOKE_BULK_IN=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_bulk.log 2>&1
grep -E "RACECHECK SUMMARY|smoke" gpurun_out/sanitizer_racecheck_bulk.log | head -4
