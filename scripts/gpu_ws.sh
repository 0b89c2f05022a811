#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== pytest ws + edge"
timeout 900 python -m pytest tests -m gpu -x -q -k "warp_specialized or non_default or smoke" 2>&1 | tail -15 | tee gpurun_out/pytest_ws.log
echo "== racecheck ws"
timeout 600 compute-sanitizer --tool racecheck --print-limit 8 python -m pytest tests/test_gpu_parity.py -q -x -k "warp_specialized and icub_like-45 and float32" > gpurun_out/racecheck_ws.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|Race reported" gpurun_out/racecheck_ws.log | head -8
timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_parity.py -q -x -k "warp_specialized and icub_like-45 and float32" > gpurun_out/memcheck_ws.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_ws.log | head -4
echo "== bench ws vs single-role"
for EXTRA in "" "--ws" "--ws --batch 8192" "--batch 65536" "--ws --batch 65536" "--ws --no-caches" "--ws --rollout 100"; do
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline $EXTRA 2>>gpurun_out/ws_err.log \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$EXTRA]', 'us/step graph=%.2f eager=%.2f'%(1e3*d['ms_per_step'],1e3*d['eager']['ms_per_step']), 'Menv/s=%.1f'%(d['value']/1e6), 'hbm_frac=%.3f'%d['roofline']['frac'], d['config']['launch'], d.get('rollout'))" \
    | tee -a gpurun_out/ws_bench.log
done
tail -3 gpurun_out/ws_err.log
