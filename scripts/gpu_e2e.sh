#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
run() {
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline "$@" 2>>gpurun_out/e2e_err.log \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); e=d['e2e']; print('[NB=%s nostep=%s | $*]' % (os.environ.get('B200SIM_E2E_BUFFERS','3'), os.environ.get('B200SIM_E2E_NOSTEP','')), 'e2e graph %.1fM (%.1f us/step) eager %.1fM' % (e['value']/1e6, d['config']['batch_per_gpu']/e['value']*1e6, e['eager_value']/1e6))"
}
mkdir -p gpurun_out
B200SIM_E2E_BUFFERS=2 run
B200SIM_E2E_BUFFERS=3 run
B200SIM_E2E_BUFFERS=4 run
B200SIM_E2E_BUFFERS=3 B200SIM_E2E_NOSTEP=1 run
B200SIM_E2E_BUFFERS=2 B200SIM_E2E_NOSTEP=1 run
B200SIM_E2E_BUFFERS=3 run --batch 16384
B200SIM_E2E_BUFFERS=3 run --batch 65536
tail -3 gpurun_out/e2e_err.log
