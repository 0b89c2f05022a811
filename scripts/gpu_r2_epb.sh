#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for opt in "" "--step-v1"; do
for E in 12 16 20 24 28 32 36; do
    timeout 300 python bench.py --batch 65536 --steps 30 --warmup 3 --no-cpu-baseline --epb $E $opt 2>>gpurun_out/epb_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('B=65536 epb=$E [$opt] us/step graph=%.2f Menv/s=%.1f frac=%.3f %s' % (1e3*d['ms_per_step'], d['value']/1e6, d['roofline']['frac'], d['config']['launch']))" | tee -a gpurun_out/epb.log
done; done
tail -3 gpurun_out/epb_err.log
