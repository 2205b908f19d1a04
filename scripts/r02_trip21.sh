#!/bin/bash
# Round-2 trip 21: TMA box cost by tensor-map rank, box size, boxes in flight and number of issuing threads.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 ./tools/tma_probe3 > gpurun_out/r02_tma_probe3.log 2>&1; echo "exit $?"; cat gpurun_out/r02_tma_probe3.log
