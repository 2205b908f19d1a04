#!/bin/bash
# Round-2 trip 18: schedule trace of update_mega_kernel (pair and single forms).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export RAFT_B200_LIB=$PWD/tools/epi_exp/mega_trace.so
timeout 200 python tools/timeline_mega.py > gpurun_out/r02_timeline_mega_pair.log 2>&1; cat gpurun_out/r02_timeline_mega_pair.log | tail -n 20
RAFT_B200_PAIR=0 timeout 200 python tools/timeline_mega.py > gpurun_out/r02_timeline_mega_single.log 2>&1; cat gpurun_out/r02_timeline_mega_single.log | tail -n 20
