#!/bin/bash
# round-2 closing pass on the final binary: full GPU parity suite, rigid cascade numbers + per-launch durations + ncu of
# the contact-QP kernel, sanitizer (racecheck) over the rigid cascade, the default bench line
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
for inp in standing random; do
  for v in "--mono" "" ; do
    echo "== $inp $v"
    python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs $inp $v 2>&1 | grep -E "counters|rigid step"
  done
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
timeout 1200 compute-sanitizer --tool racecheck --print-limit 5 python scripts/san_rigid.py > gpurun_out/sanitizer_rigid_racecheck.log 2>&1
echo "rigid racecheck: $(grep -E 'RACECHECK SUMMARY' gpurun_out/sanitizer_rigid_racecheck.log | head -1)"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python scripts/san_rigid.py > gpurun_out/sanitizer_rigid_memcheck.log 2>&1
echo "rigid memcheck: $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer_rigid_memcheck.log | head -1)"
echo "== rigid + relaxed-rigid legs"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --config3 2>gpurun_out/c3_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k in ('config3_rigid','relaxed_rigid'):
    c=d[k]
    for lab in ('random','standing'): print(k, lab, 'ms/step %.3f' % c[lab]['ms_per_step'], 'env-steps/s %.3e' % c[lab]['value'])"
echo "== default bench line"
timeout 1200 python bench.py 2>gpurun_out/bench_err.log > gpurun_out/bench_default.json
tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_default.json"))
print("value %.4e ms/step %.5f frac %.4f e2e %.3e" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))
c=d.get("config3_rigid") or {}
for k in ("random","standing"):
    if k in c: print("config3", k, "%.3f ms %.3e" % (c[k]["ms_per_step"], c[k]["value"]))
print("large", [(l["batch"], round(l["ms_per_step"]*1e3,1), round(l["roofline"]["frac"],3)) for l in d.get("large_batch",[])])
PY
