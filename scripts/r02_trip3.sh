#!/bin/bash
# Round-2 trip 3: cleaned-up library (dead variants removed, PDL default), training step, lookup v2, configs 3 / 4, encoder chunking A/B.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
bench() {  # name, args, env...
  local name=$1 bargs=$2; shift 2
  env "$@" timeout 400 python bench.py $bargs > gpurun_out/r02_b3_$name.json 2> gpurun_out/r02_b3_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/r02_b3_{name}.json'))
    extra = ''
    if 'roofline_corr_lookup' in d: extra = f"corr/lookup ms {d['roofline_corr_lookup']['ms']} mega {d['roofline']['ms_per_launch']:.4f} ms {d['roofline']['achieved']:.1f} TF"
    print(f"bench {name:<14} {d['value']:8.2f} pairs/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.2f}  {extra} {d['config'].get('allreduce_ms','')} {d.get('loss_first_last','')}")
except Exception as e:
    print(f'bench {name}: FAILED ({e})'); print(open(f'gpurun_out/r02_b3_{name}.err').read()[-2500:])
PY
}
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_t3.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t3.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t3.log | head -20
bench default "--steps 10 --warmup 3 --quick" A=1
bench nopdl "--steps 10 --warmup 3 --quick" RAFT_B200_PDL=0
bench encimg2 "--steps 10 --warmup 3 --quick" RAFT_B200_ENC_IMAGES=2
bench encimg4 "--steps 10 --warmup 3 --quick" RAFT_B200_ENC_IMAGES=4
bench sintel "--config sintel --steps 5 --warmup 3 --quick" A=1
bench train "--config train --steps 5 --warmup 3 --quick" A=1
