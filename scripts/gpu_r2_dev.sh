#!/bin/bash
# round-2 development call: parity of the second-generation step kernel + A/B timing against the first generation
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
nvidia-smi -L | head -1; nproc
echo "== pytest (soft-step parity + goldens)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_goldens.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -15 | cut -c1-3000 | tee gpurun_out/pytest_gpu_dev.log
echo "== A/B"
for B in 4096 16384 65536; do
  for opt in "--step-v1" "" "--no-bulk-in" "--lanes 16"; do
    timeout 300 python bench.py --batch $B --steps 60 --warmup 5 --no-cpu-baseline $opt 2>>gpurun_out/ab_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('B=$B [$opt] us/step graph=%.2f eager=%.2f Menv/s=%.1f frac=%.3f e2e=%.1f %s' % (1e3*d['ms_per_step'], 1e3*d['eager']['ms_per_step'], d['value']/1e6, d['roofline']['frac'], d['e2e']['value']/1e6, d['config']['launch']))" | tee -a gpurun_out/ab.log
  done
done
tail -5 gpurun_out/ab_err.log
