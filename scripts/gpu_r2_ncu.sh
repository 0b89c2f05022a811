#!/bin/bash
# ncu --set full of the step kernel: $1 = tag, rest = bench.py options
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
tag=$1; shift
timeout 900 ncu --profile-from-start off --set full --clock-control none --cache-control none --import-source on -k regex:step -c 1 -f -o gpurun_out/prof_$tag \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --profile "$@" > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-300
ls -la gpurun_out/prof_$tag.ncu-rep
