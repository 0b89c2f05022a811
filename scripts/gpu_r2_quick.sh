#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_rigid.py -m gpu -q 2>&1 | tail -3
for v in "" "--dtype f64"; do python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs standing $v 2>&1 | grep -E "rigid step"; done
python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs random 2>&1 | grep -E "rigid step"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/split_launches_standing.csv \
    python scripts/rigid_profile.py --batch 16384 --steps 1 --inputs standing > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/split_launches_standing.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows:
    print("  %-62s %s %s" % (r[4][:60], r[-1], r[-2]))
PY
