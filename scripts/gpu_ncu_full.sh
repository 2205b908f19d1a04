#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# skip the weight-prep + first forward's conv_tc launches (~62 per forward: 2x19 encoder + 4 corr + 12x11 loop = 174), land in forward 2's loop
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 230 -c 12 -o gpurun_out/prof_conv_tc python scripts/profile_loop.py f16x2 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 300 ncu --set full --clock-control none -k regex:corr_lookup_kernel -s 14 -c 1 -o gpurun_out/prof_lookup python scripts/profile_loop.py f16x2 2 > gpurun_out/ncu_full2.log 2>&1; echo "ncu lookup exit $?"
ls -la gpurun_out/*.ncu-rep
