#!/bin/bash
# multi-GPU check of the bench contract: torchrun launch as the driver does it, default weak scaling and BASELINE configs[3]
# (65 536 environments global), plus the reference arm under torchrun (rank 0 prints, the others exit 0)
set -u
N=${1:-4}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
echo "== weak scaling, 4096 per GPU"
run --steps 20 --warmup 3 2>gpurun_out/multi_err.log | tee gpurun_out/bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('n_gpus', d['n_gpus'], 'us/step %.2f value %.3e e2e %.3e copy-only %s numa %s' % (1e3*d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e'].get('copy_only_value'), d['e2e'].get('numa')))"
echo "== configs[3]: 65536 global"
run --steps 20 --warmup 3 --global-batch 65536 2>>gpurun_out/multi_err.log | tee gpurun_out/bench_n${N}_g65536.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('n_gpus', d['n_gpus'], 'us/step %.2f value %.3e frac %.3f e2e %.3e' % (1e3*d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value']), d['config'].get('global_batch'))"
echo "== reference arm under torchrun"
run --impl reference --steps 3 --warmup 1 2>>gpurun_out/multi_err.log | cut -c1-400
tail -3 gpurun_out/multi_err.log
