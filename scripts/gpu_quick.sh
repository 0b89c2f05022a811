#!/bin/bash
# quick GPU round: full parity suite + default bench (+ optional extra commands via $EXTRA)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
nvidia-smi -L | head -2
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== bench (default)"
timeout 600 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_default.json | cut -c1-2500
tail -5 gpurun_out/bench_err.log
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_reference.json | cut -c1-600
if [ -n "${EXTRA:-}" ]; then echo "== extra: $EXTRA"; bash -c "$EXTRA"; fi
