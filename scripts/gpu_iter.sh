#!/bin/bash
# one optimisation iteration on the GPU: parity suite, default bench, phase clocks, racecheck of the smoke step
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | cut -c1-2000 | tee gpurun_out/pytest_gpu.log
run() {
  timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline "$@" 2>>gpurun_out/ab_err.log \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$*]', 'us/step graph=%.2f eager=%.2f'%(1e3*d['ms_per_step'],1e3*d['eager']['ms_per_step']), 'Menv/s=%.1f'%(d['value']/1e6), 'frac=%.3f'%d['roofline']['frac'], 'e2e=%.1fM'%(d['e2e']['value']/1e6), d['config']['launch'])" \
    | tee -a gpurun_out/ab.log
}
run
run --batch 8192
run --batch 65536
run --dtype f64
python scripts/phase_clocks.py 2>&1 | tee gpurun_out/phase_clocks.log
echo "== racecheck / memcheck (smoke)"
for tool in racecheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke|hazard" gpurun_out/sanitizer_$tool.log | head -8
done
tail -3 gpurun_out/ab_err.log
