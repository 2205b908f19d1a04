#!/bin/bash
# Round-2 multi-GPU trip (gpurun --gpus 4): inference weak scaling (chairs, sintel) and the data-parallel training step.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8
run() {  # name, nproc, bench args
  local name=$1 n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" > gpurun_out/r02_n_$name.json 2> gpurun_out/r02_n_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/r02_n_{name}.json'))
    print(f"{name:<12} n={d['n_gpus']} {d['value']:9.2f} pairs/s {d['ms_per_step']:.3f} ms/step e2e {d['e2e']['value']:.2f} allreduce_ms {d['config'].get('allreduce_ms')} loss {d.get('loss_first_last')}")
except Exception as e:
    print(f'{name}: FAILED ({e})'); print(open(f'gpurun_out/r02_n_{name}.err').read()[-3000:])
PY
}
run chairs4 4 --steps 10 --warmup 3 --quick
run sintel4 4 --config sintel --steps 5 --warmup 3 --quick
run train4 4 --config train --steps 5 --warmup 3 --quick
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 > gpurun_out/r02_n_ref4.json 2> gpurun_out/r02_n_ref4.err; echo "reference arm under torchrun exit $?: $(head -c 300 gpurun_out/r02_n_ref4.json)"
