#!/bin/bash
# Round-2 multi-GPU trip (gpurun --gpus 8): inference weak scaling on a full box.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 --quick > gpurun_out/r02_n_chairs8.json 2> gpurun_out/r02_n_chairs8.err; echo "exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r02_n_chairs8.json')); print('chairs8', d['n_gpus'], round(d['value'],1), 'pairs/s', d['ms_per_step'], 'ms/step e2e', round(d['e2e']['value'],1))" || tail -n 20 gpurun_out/r02_n_chairs8.err
