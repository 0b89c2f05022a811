#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3
for B in 4096 16384 65536; do
  for opt in "--step-v1" ""; do
    timeout 300 python bench.py --batch $B --steps 60 --warmup 5 --no-cpu-baseline --no-extras $opt 2>>gpurun_out/ab_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('B=$B [$opt] us/step graph=%.2f eager=%.2f Menv/s=%.1f frac=%.3f e2e=%.1f %s' % (1e3*d['ms_per_step'], 1e3*d['eager']['ms_per_step'], d['value']/1e6, d['roofline']['frac'], d['e2e']['value']/1e6, d['config']['launch']))" | tee -a gpurun_out/ab3.log
  done
done
python scripts/phase_clocks.py 2>&1 | head -19
tail -3 gpurun_out/ab_err.log
