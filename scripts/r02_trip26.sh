#!/bin/bash
# Round-2 trip 26: which part of the mainloop bounds the narrow layers (tools/tc_exp.sh).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 bash tools/tc_exp.sh run > gpurun_out/r02_tc_exp.log 2>&1; cat gpurun_out/r02_tc_exp.log
