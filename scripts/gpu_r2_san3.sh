#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --print-limit 5 python scripts/san_rigid.py > gpurun_out/sanitizer_rigid_racecheck.log 2>&1
echo "rigid racecheck: $(grep -E 'RACECHECK SUMMARY' gpurun_out/sanitizer_rigid_racecheck.log | head -1)"
grep -h "rigid split cascade\|Race reported" gpurun_out/sanitizer_rigid_racecheck.log | head
timeout 1200 python -m pytest tests/test_gpu_rigid.py tests/test_gpu_reference_goldens.py -m gpu -q -k "rigid" 2>&1 | tail -3
python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs standing 2>&1 | grep -E "rigid step"
python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs random 2>&1 | grep -E "rigid step"
