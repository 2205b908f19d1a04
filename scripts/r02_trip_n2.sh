#!/bin/bash
# Round-2 multi-GPU trip (gpurun --gpus 2): inference weak scaling and the data-parallel training step over NCCL.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv
run() {  # name, nproc, bench args
  local name=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/r02_n_$name.json 2> gpurun_out/r02_n_$name.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" > gpurun_out/r02_n_$name.json 2> gpurun_out/r02_n_$name.err; fi
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
run chairs1 1 --steps 10 --warmup 3 --quick
run chairs2 2 --steps 10 --warmup 3 --quick
run train1 1 --config train --steps 5 --warmup 3 --quick
run train2 2 --config train --steps 5 --warmup 3 --quick
