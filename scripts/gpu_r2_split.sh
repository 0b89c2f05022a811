#!/bin/bash
# split rigid cascade (assemble / solve / resume, level 2 on a side stream): parity, A/B against the monolithic kernel,
# per-kernel durations, ncu --set full of the contact-QP kernel
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 1200 python -m pytest tests/test_gpu_rigid.py tests/test_gpu_reference_goldens.py tests/test_gpu_edge_cases.py -m gpu -q 2>&1 | tail -6
for inp in standing random; do
  for v in "--mono" "" ; do
    echo "== $inp $v"
    python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs $inp $v 2>&1 | grep -E "counters|rigid step"
  done
done
for inp in standing random; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/split_launches_${inp}.csv \
    python scripts/rigid_profile.py --batch 16384 --steps 1 --inputs $inp > /dev/null 2>&1
  echo "== launches $inp"
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/split_launches_${inp}.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows:
    print("  %-62s %s %s" % (r[4][:60], r[-1], r[-2]))
PY
done
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rigid_qp_kernel -c 1 -f -o gpurun_out/prof_rigid_qp \
  python scripts/rigid_profile.py --batch 16384 --steps 1 > gpurun_out/ncu_rigid_qp.log 2>&1
ls -la gpurun_out/prof_rigid_qp.ncu-rep
