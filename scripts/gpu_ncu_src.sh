#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
# forward 2's loop: conv_tc launch index: prep none (conv_tc only in forwards): per forward 19+19(enc)... use the first loop layers of forward 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 230 -c 2 -o gpurun_out/prof_src python scripts/profile_loop.py f16x2 2 > gpurun_out/ncu_src.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/prof_src.ncu-rep
