cd "${GRAFT_REPO_ROOT:-/root/repo}"
for L in 8 16; do for B in 4096 16384; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --model ergocub_like --lanes $L --batch $B 2>>gpurun_out/ergo_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ergocub soft lanes=$L B=$B us/step=%.2f Menv/s=%.1f frac=%.3f'%(1e3*d['ms_per_step'], d['value']/1e6, d['roofline']['frac']), d['config']['launch'])"
done; done; tail -2 gpurun_out/ergo_err.log
