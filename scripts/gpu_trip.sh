#!/bin/bash
# One GPU-box visit: isolated pytest groups (a CUDA fault in one must not poison the next), smoke, bench.
# Logs go to gpurun_out/ (merged back by gpurun).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('cores', os.cpu_count())" >> gpurun_out/gpu.txt
python -c "import tensorflow" >> gpurun_out/gpu.txt 2>&1 || echo "tensorflow: not installed" >> gpurun_out/gpu.txt
run() {  # name, timeout, pytest -k expression
  echo "=== $1" | tee -a gpurun_out/summary.txt
  timeout "$2" python -m pytest tests -m gpu -q -k "$3" -p no:cacheprovider > "gpurun_out/$1.log" 2>&1
  echo "exit $? : $(tail -n 1 gpurun_out/$1.log)" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run t1_simt     600 "fp32 or bit_exact or sampler or coords_grid or upsample_convex or correlation_method"
run t2_tc_corr  300 "f16x2 and corr_pyramid"
run t3_tc_upd   300 "f16x2 and update_block"
run t4_tc_model 400 "f16x2 and (small_raft or reference_test_shape) or api_contract"
run t5_tc_full  600 "f16x2 and (448x512 or full_size)"
echo "=== smoke" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/smoke.log | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
for prec in fp32 f16x2; do
  echo "=== bench $prec" | tee -a gpurun_out/summary.txt
  timeout 600 python bench.py --steps 5 --warmup 3 --precision $prec > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err
  echo "exit $? : $(head -c 600 gpurun_out/bench_$prec.json)" | tee -a gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
