#!/bin/bash
# A/B run of the experiment knobs (scripts/README.md): for each, the relevant GPU parity tests with the knob on, then
# `bench.py --quick`.  One gpurun call, about 4-5 minutes.  Usage: gpurun --timeout 900 -- 'bash scripts/gpu_ab.sh'
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
bench() {  # name, env...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/ab_{name}.json'))
    print(f"bench {name:<14} {d['value']:8.1f} pairs/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.1f}  corr/lookup ms {d['roofline_corr_lookup']['ms']}")
except Exception as e:
    print(f'bench {name}: FAILED ({e})')
PY
}
check() {  # name, pytest -k expression, env...
  local name=$1 k=$2; shift 2
  env "$@" timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "$k" > gpurun_out/ab_test_$name.log 2>&1
  echo "tests $name: exit $? : $(tail -n 1 gpurun_out/ab_test_$name.log)"
}
bench baseline A=1
check fused_stats "encoder or reference_test_shape or small_raft" RAFT_B200_FUSED_STATS=1
bench fused_stats RAFT_B200_FUSED_STATS=1
check enc_swap "encoder or reference_test_shape or small_raft" RAFT_B200_ENC_SWAP=2
bench enc_swap RAFT_B200_ENC_SWAP=1
bench swap_and_stats RAFT_B200_ENC_SWAP=1 RAFT_B200_FUSED_STATS=1
check pdl "update_block or reference_test_shape or corr_pyramid" RAFT_B200_PDL=1
bench pdl RAFT_B200_PDL=1
bench lookup_gather RAFT_B200_LOOKUP_GATHER=1
check fh2_simt "update_block or reference_test_shape or small_raft" RAFT_B200_FH2_SIMT=1
bench fh2_simt RAFT_B200_FH2_SIMT=1
check two_streams "reference_test_shape or model_api or other_resolutions" RAFT_B200_TWO_STREAMS=1
bench two_streams RAFT_B200_TWO_STREAMS=1
check corr_tma "corr or pyramid or reference_test_shape or small_raft" RAFT_B200_CORR_TMA_STORE=1
bench corr_tma RAFT_B200_CORR_TMA_STORE=1
timeout 200 python bench.py --steps 10 --warmup 3 --quick --pipeline > gpurun_out/ab_pipeline.json 2> gpurun_out/ab_pipeline.err; python -c "import json; d=json.load(open('gpurun_out/ab_pipeline.json')); print('bench pipeline e2e', d['e2e'])"
bench all RAFT_B200_TWO_STREAMS=1 RAFT_B200_FH2_SIMT=1 RAFT_B200_CORR_TMA_STORE=1 RAFT_B200_ENC_SWAP=1 RAFT_B200_FUSED_STATS=1 RAFT_B200_PDL=1
