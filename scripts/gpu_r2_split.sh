#!/bin/bash
# split rigid cascade (assemble / solve / resume, level 2 on a side stream): parity, then A/B against the monolithic kernel
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python scripts/debug_split.py 2>&1 | grep -A4 "split vs mono"
timeout 1200 python -m pytest tests/test_gpu_rigid.py tests/test_gpu_reference_goldens.py tests/test_gpu_edge_cases.py -m gpu -q 2>&1 | tail -12
for inp in standing random; do
  for v in "--mono" "" ; do
    echo "== $inp $v"
    python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs $inp $v 2>&1 | grep -E "counters|rigid step"
  done
done
