#!/bin/bash
# Round-2 trip 13: norm_stats_kernel register cap (the fused finalisation raised it to 80 registers -> 3 blocks per SM -> a
# second wave, +52 us per launch); cost breakdown of the convolution epilogues (tools/epi_exp.sh).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/parity_probe.py regcap 2>&1 | tail -n 1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "encoder or model or parity" > gpurun_out/r02_t13.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t13.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t13.log | head -12
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b13.json 2>gpurun_out/r02_b13.err
python -c "
import json; d=json.load(open('gpurun_out/r02_b13.json')); print('bench', round(d['value'],1), 'pairs/s e2e', round(d['e2e']['value'],1), d['roofline_corr_lookup']['ms'], 'mega ms', round(d['roofline']['ms_per_launch'],4))"
timeout 900 bash tools/epi_exp.sh run > gpurun_out/r02_epi_exp.log 2>&1; cat gpurun_out/r02_epi_exp.log
