#!/bin/bash
# Round-2 trip 39: the bench line of the final tree.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench.json')); print(d['metric'], d['value'], d['e2e']['value'], d['parity'].get('max_abs'), d['roofline']['frac'], d['roofline']['shared_memory']['frac'])"
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; echo "reference exit $?"
