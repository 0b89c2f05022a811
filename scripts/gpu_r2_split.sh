#!/bin/bash
# split rigid level (assemble / solve / resume): parity, then A/B against the monolithic kernel and register variants
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 1200 python -m pytest tests/test_gpu_rigid.py tests/test_gpu_reference_goldens.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -4
for inp in standing random; do
  for v in "--mono" "" ; do
    echo "== $inp $v"
    python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs $inp $v 2>&1 | grep -E "counters|rigid step"
  done
  for mb in 8; do
    echo "== $inp MINB=$mb"
    B200SIM_QP_MINB=$mb python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs $inp 2>&1 | grep -E "rigid step"
  done
done
# per-kernel durations (serialised by ncu): where the step's time goes in both modes
for inp in standing random; do
  for v in "--mono" "" ; do
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/split_launches_${inp}${v}.csv \
      python scripts/rigid_profile.py --batch 16384 --steps 1 --inputs $inp $v > /dev/null 2>&1
    echo "== launches $inp $v"
    python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/split_launches_${inp}${v}.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows:
    name=r[4][:60]; val=r[-1]; unit=r[-2]
    print("  %-62s %s %s" % (name, val, unit))
PY
  done
done
