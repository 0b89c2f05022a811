#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "== output-path diagnostics (B=4096, graph timing)"
for EXTRA in "" "--out-ring 1" "--skip-cache X" "--skip-cache H" "--skip-cache V" "--skip-cache XHV" "--skip-cache XHVB" "--no-caches" "--no-tma --skip-cache HV"; do
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline $EXTRA 2>>gpurun_out/diag_err.log \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$EXTRA]', 'us/step graph=%.2f eager=%.2f'%(1e3*d['ms_per_step'],1e3*d['eager']['ms_per_step']))" \
    | tee -a gpurun_out/diag.log
done
tail -3 gpurun_out/diag_err.log
