#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 1200 python -m pytest tests/test_gpu_rigid.py tests/test_gpu_reference_goldens.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --config3 2>gpurun_out/c3_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k in ('config3_rigid','relaxed_rigid'):
    c=d[k]
    for lab in ('random','standing'): print(k, lab, 'ms/step %.3f' % c[lab]['ms_per_step'], 'env-steps/s %.3e' % c[lab]['value'], 'act', c[lab]['mean_active_points'])"
tail -3 gpurun_out/c3_err.log
python scripts/rigid_profile.py 2>&1 | tail -16
