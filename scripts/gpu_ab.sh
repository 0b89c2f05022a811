#!/bin/bash
# A/B timing of step-kernel variants (graph timing, B=4096 and larger) + the parity suite
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
run() {
  timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline "$@" 2>>gpurun_out/ab_err.log \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$*]', 'us/step graph=%.2f eager=%.2f'%(1e3*d['ms_per_step'],1e3*d['eager']['ms_per_step']), 'Menv/s=%.1f'%(d['value']/1e6), 'frac=%.3f'%d['roofline']['frac'], 'e2e=%.1fM'%(d['e2e']['value']/1e6), d['config']['launch'])" \
    | tee -a gpurun_out/ab.log
}
for extra in ${AB_VARIANTS:-"" "--generic-kernel" "--lanes 16" "--lanes 16 --generic-kernel" "--batch 8192" "--batch 8192 --generic-kernel" "--batch 65536" "--batch 65536 --generic-kernel" "--no-caches" "--no-caches --generic-kernel"}; do
  run $extra
done
tail -3 gpurun_out/ab_err.log
