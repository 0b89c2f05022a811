#!/bin/bash
# ncu --set full of the contact-QP kernel of the split rigid cascade (standing ErgoCub-like inputs)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rigid_qp_kernel -c 1 -f -o gpurun_out/prof_rigid_qp \
  python scripts/rigid_profile.py --batch 16384 --steps 1 > gpurun_out/ncu_rigid_qp.log 2>&1
tail -3 gpurun_out/ncu_rigid_qp.log | cut -c1-300
ls -la gpurun_out/prof_rigid_qp.ncu-rep
