#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | cut -c1-2000 | tee gpurun_out/pytest_gpu.log
run() {
  timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline "$@" 2>>gpurun_out/ab_err.log \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('[$*]', 'us/step graph=%.2f eager=%.2f'%(1e3*d['ms_per_step'],1e3*d['eager']['ms_per_step']), 'Menv/s=%.1f'%(d['value']/1e6), 'e2e=%.1fM/%.1fM' % (d['e2e']['value']/1e6, d['e2e']['eager_value']/1e6))" \
    | tee -a gpurun_out/ab.log
}
for rep in 1 2; do
run
run --no-pdl
done
run --batch 8192
run --batch 8192 --no-pdl
run --batch 65536
tail -3 gpurun_out/ab_err.log
