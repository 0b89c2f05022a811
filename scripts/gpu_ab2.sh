#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
run() {
  timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline "$@" 2>>gpurun_out/ab_err.log \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('[%s | $*]' % os.environ.get('B200SIM_LIB','default')[-12:], 'us/step graph=%.2f eager=%.2f'%(1e3*d['ms_per_step'],1e3*d['eager']['ms_per_step']), 'Menv/s=%.1f'%(d['value']/1e6), d['config']['launch'])" \
    | tee -a gpurun_out/ab.log
}
for rep in 1 2; do
run
run --no-bulk-in
B200SIM_LIB=$PWD/jaxsim_b200/csrc/libb200sim_t288.so run
B200SIM_LIB=$PWD/jaxsim_b200/csrc/libb200sim_t288.so run --no-bulk-in
done
run --batch 8192
run --batch 8192 --no-bulk-in
B200SIM_LIB=$PWD/jaxsim_b200/csrc/libb200sim_t288.so run --batch 8192
B200SIM_LIB=$PWD/jaxsim_b200/csrc/libb200sim_t288.so run --batch 8192 --no-bulk-in
python scripts/phase_clocks.py
B200SIM_LIB=$PWD/jaxsim_b200/csrc/libb200sim_t288.so python scripts/phase_clocks.py
tail -3 gpurun_out/ab_err.log
