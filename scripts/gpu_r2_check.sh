#!/bin/bash
# round-2 check: full GPU parity suite + the default bench line (with its bounded extra legs)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
nvidia-smi -L | head -1; nproc
echo "== pytest -m gpu"
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | cut -c1-3000 | tee gpurun_out/pytest_gpu.log
echo "== bench (default)"
timeout 900 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_default.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('us/step %.2f value %.3e frac %.3f e2e %.3e replays %s launches %s' % (1e3*d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('replays'), d.get('gpu_launches')))
for k in ('large_batch','no_pdl','in_contact','config3_rigid','config5_jvp','cpu_baseline','compute','clocks'): print(k, json.dumps(d.get(k))[:900])"
tail -5 gpurun_out/bench_err.log
