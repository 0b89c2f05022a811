#!/bin/bash
# ncu --set full of the RESUME launch of the split rigid level (third rigid_step_kernel launch of a step), standing inputs
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rigid_step_kernel -s 2 -c 1 -f -o gpurun_out/prof_rigid_resume \
  python scripts/rigid_profile.py --batch 16384 --steps 1 > gpurun_out/ncu_rigid_resume.log 2>&1
ls -la gpurun_out/prof_rigid_resume.ncu-rep
