#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/summary.txt
run() {
  echo "=== $1" | tee -a gpurun_out/summary.txt
  timeout "$2" python -m pytest tests -m gpu -q -s -k "$3" -p no:cacheprovider > "gpurun_out/$1.log" 2>&1
  echo "exit $? : $(tail -n 1 gpurun_out/$1.log)" | tee -a gpurun_out/summary.txt
}
run s1_stages 600 "stages"
run s2_full   600 "448x512 or ragged"
for prec in f16x2 fp32; do
  echo "=== bench $prec" | tee -a gpurun_out/summary.txt
  timeout 480 python bench.py --steps 5 --warmup 3 --precision $prec > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err
  echo "exit $? : $(head -c 400 gpurun_out/bench_$prec.json)" | tee -a gpurun_out/summary.txt
  tail -n 30 gpurun_out/bench_$prec.err
done
grep -hE "max-abs|encoder|pyramid level|update block|upsample iteration" gpurun_out/s1_stages.log gpurun_out/s2_full.log | head -60
