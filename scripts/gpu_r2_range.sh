#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 ncu --replay-mode range --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv \
  --log-file gpurun_out/range_B4096.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-graph --no-extras --profile > gpurun_out/range_B4096.log 2>&1
tail -4 gpurun_out/range_B4096.csv | cut -c1-400
tail -3 gpurun_out/range_B4096.log | cut -c1-300
for NB in 2 3 4; do
B200SIM_E2E_BUFFERS=$NB timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras 2>>gpurun_out/ab_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('NB=$NB e2e %.1f (%s) eager %.1f copy-only %.1f numa %s' % (e['value']/1e6, e['launch'][:40], e['eager_value']/1e6, e['copy_only_value']/1e6, e['numa']))"
done
nvidia-smi topo -m | head -8; lscpu | grep -E "NUMA|Socket|Model name" | head -8; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; python -c "import os; print(sorted(os.sched_getaffinity(0)))"
