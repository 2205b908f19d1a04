#!/bin/bash
# Round-2 trip 23: tcgen05.mma issue-rate probe.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 60 ./tools/mma_probe > gpurun_out/r02_mma_probe.log 2>&1; echo "exit $?"; cat gpurun_out/r02_mma_probe.log
