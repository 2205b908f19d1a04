#!/bin/bash
# Round-2 trip 1: full GPU suite (incl. the new graph / last_only / batch-4 tests), A/B of every experiment knob
# (parity tests with the knob on + bench --quick), config-5 sweep of the round-1 kernels.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_t_all.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t_all.log)"
grep -hE "FAILED|Error" gpurun_out/r02_t_all.log | head -10
timeout 300 python tools/corr_sweep.py > gpurun_out/r02_corr_sweep_r01kernels.log 2>&1; echo "sweep exit $?"; cat gpurun_out/r02_corr_sweep_r01kernels.log
bash scripts/gpu_ab.sh 2>&1 | tee gpurun_out/r02_ab.log
