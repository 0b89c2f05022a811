#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3
for B in 4096 65536; do
    timeout 300 python bench.py --batch $B --steps 60 --warmup 5 --no-cpu-baseline --no-extras 2>>gpurun_out/ab_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('B=$B us/step graph=%.2f eager=%.2f Menv/s=%.1f frac=%.3f e2e=%.1f %s' % (1e3*d['ms_per_step'], 1e3*d['eager']['ms_per_step'], d['value']/1e6, d['roofline']['frac'], d['e2e']['value']/1e6, d['config']['launch']))"
done
echo "== sanitizers (smoke)"
for tool in racecheck memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke" gpurun_out/sanitizer_$tool.log | head -4
done
echo "== range replay: DRAM traffic of 40 back-to-back launches"
timeout 600 ncu --replay-mode range --profile-from-start off --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv \
  --log-file gpurun_out/range_B4096.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-graph --no-extras --profile > gpurun_out/range_B4096.log 2>&1
tail -4 gpurun_out/range_B4096.csv | cut -c1-300
tail -3 gpurun_out/range_B4096.log | cut -c1-300
