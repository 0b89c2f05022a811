#!/bin/bash
# compute-sanitizer over the soft step (smoke) and the split rigid cascade, final binary of round 2
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "smoke $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | head -1)"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python scripts/san_rigid.py > gpurun_out/sanitizer_rigid_$tool.log 2>&1
  echo "rigid $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_rigid_$tool.log | head -1) | $(grep -c 'rigid split cascade' gpurun_out/sanitizer_rigid_$tool.log) result lines"
done
grep -h "rigid split cascade" gpurun_out/sanitizer_rigid_memcheck.log
