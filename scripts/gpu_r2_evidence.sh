#!/bin/bash
# round-2 evidence for the FINAL binary: counters (steady-state DRAM traffic + flops), ncu launch list, ncu --set full
# (batch 4096 and 65536, warm caches), phase clocks, sanitizers
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
M="dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,gpu__time_duration.sum"
echo "== counters"
for B in 4096 65536; do
timeout 600 ncu --profile-from-start off --cache-control none --clock-control none -k regex:step -c 40 --csv --metrics $M \
  --log-file gpurun_out/step_counters_B$B.csv python bench.py --batch $B --steps 40 --warmup 5 --no-cpu-baseline --no-graph --no-extras --profile > gpurun_out/counters_B$B.log 2>&1
tail -1 gpurun_out/step_counters_B$B.csv | cut -c1-200
done
echo "== ncu launch list (default bench command, eager launches)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph --no-extras --profile > gpurun_out/ncu_launch_bench.log 2>&1
grep -c step gpurun_out/launches.csv
echo "== ncu full"
bash scripts/gpu_r2_ncu.sh final_b4k --batch 4096 --no-extras
bash scripts/gpu_r2_ncu.sh final_b65k --batch 65536 --no-extras
echo "== phase clocks"
python scripts/phase_clocks.py 2>&1 | tee gpurun_out/phase_clocks.log
python scripts/phase_clocks.py --ring 12 2>&1 | tail -2 | tee -a gpurun_out/phase_clocks.log
python scripts/phase_clocks.py --step-v1 2>&1 | head -20 | tee gpurun_out/phase_clocks_v1.log
echo "== sanitizers (smoke)"
for tool in racecheck memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke" gpurun_out/sanitizer_$tool.log | head -4
done
