#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rigid.py tests/test_gpu_reference_goldens.py tests/test_gpu_edge_cases.py -m gpu -q 2>&1 | tail -8
for inp in standing random; do
  for v in "--mono" "" ; do
    echo "== $inp $v"
    python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs $inp $v 2>&1 | grep -E "counters|rigid step"
  done
done
python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs standing --dtype f64 2>&1 | grep -E "counters|rigid step"
