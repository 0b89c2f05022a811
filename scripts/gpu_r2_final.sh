#!/bin/bash
# last pass on the final binary: new tests, sanitizers over smoke(), the ErgoCub-like soft line for the record
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_autodiff.py tests/test_gpu_parity.py -q -m gpu -k "full_batch_properties or rollout_records" 2>&1 | tail -4
echo "== ergocub_like soft"
timeout 600 python bench.py --model ergocub_like --no-extras --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_ergocub_soft.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('us/step %.2f value %.3e frac %.3f' % (1e3*d['ms_per_step'], d['value'], d['roofline']['frac']), d['config'].get('launch'))"
echo "== sanitizers (smoke)"
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke" gpurun_out/sanitizer_$tool.log | head -4
done
