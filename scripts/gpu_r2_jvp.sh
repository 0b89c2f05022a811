#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_autodiff.py tests/test_gpu_parity.py -q -m gpu -k "jvp or jacobian or vjp or update_link" -x 2>&1 | tail -5
timeout 600 python bench.py --jvp --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_jvp.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(json.dumps(d.get('config5_jvp'))[:900]); print('rollout', json.dumps(d.get('rollout'))[:400])"
tail -3 gpurun_out/bench_err.log
