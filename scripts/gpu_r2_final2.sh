#!/bin/bash
# round-2 closing pass: full GPU parity suite, tolerance A/B of the contact QP, the default bench line
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== contact QP stopped at 1e-6 for float32 data (default 1e-8): parity + time"
B200SIM_QP_TOL_F32=1e-6 timeout 900 python -m pytest tests/test_gpu_rigid.py tests/test_gpu_reference_goldens.py -m gpu -q -k "rigid and not relaxed" 2>&1 | tail -3
B200SIM_QP_TOL_F32=1e-6 python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs standing 2>&1 | grep -E "counters|rigid step"
python scripts/rigid_profile.py --batch 16384 --steps 5 --inputs standing 2>&1 | grep -E "counters|rigid step"
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
print("jvp", {k: d["config5_jvp"][k] for k in ("ms_per_jvp","ms_full_jacobian","ms_vjp") if k in d.get("config5_jvp",{})})
PY
