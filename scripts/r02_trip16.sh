#!/bin/bash
# Round-2 trip 16: CTA-pair (cta_group::2) probe; A/B of the update_mega_kernel producer variants on one box.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 60 ./tools/pair_probe > gpurun_out/r02_pair_probe.log 2>&1; echo "pair probe exit $?"; cat gpurun_out/r02_pair_probe.log
timeout 900 bash tools/mega_ab.sh run > gpurun_out/r02_mega_ab.log 2>&1; cat gpurun_out/r02_mega_ab.log
