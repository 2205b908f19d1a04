#!/bin/bash
# Round-2 trip 17: update_mega_kernel<true> (CTA pairs, tcgen05 cta_group::2) against the single-CTA form.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 60 ./tools/pair_probe > gpurun_out/r02_pair_probe.log 2>&1; echo "pair probe exit $?"; cat gpurun_out/r02_pair_probe.log
RAFT_B200_PAIR=0 timeout 200 python scripts/parity_probe.py single 2>&1 | tail -n 1
timeout 200 python scripts/parity_probe.py pair 2>&1 | tail -n 3
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_t17.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t17.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t17.log | head -12
for rep in 1 2; do
for pair in 0 1; do
RAFT_B200_PAIR=$pair timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b17_$pair.json 2>gpurun_out/r02_b17_$pair.err
python -c "
import json; d=json.load(open('gpurun_out/r02_b17_$pair.json')); print('pair=$pair bench', round(d['value'],1), 'pairs/s e2e', round(d['e2e']['value'],1), 'mega ms', round(d['roofline']['ms_per_launch'],4), d['parity'].get('max_abs'))"
done
done
