#!/bin/bash
# What bounds the mainloop of the narrow layers?  Builds of the library with parts of the per-layer kernel's mainloop compiled
# out (RAFT_TC_EXP bit 0: no MMAs, only the commits; 2: activation boxes only; 4: weight boxes only) and the per-chunk
# timeline of update-block layers 9 (flow-head conv2, N = 16), 3 (conv, N = 128), 4 (GRU zr, N = 256).
#   build (no GPU):  bash tools/tc_exp.sh build        run (GPU box):  bash tools/tc_exp.sh run
set -u
cd "$(dirname "$0")/.."
VARIANTS="0 1 2 3 4 5"
mkdir -p tools/epi_exp
if [ "${1:-run}" = build ]; then
  for v in $VARIANTS; do
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DRAFT_TC_EXP=$v \
      tf_raft_b200/csrc/api.cu -o tools/epi_exp/tc_$v.so &
  done
  wait
  ls -la tools/epi_exp/
  exit 0
fi
for v in $VARIANTS; do
  for layer in 9 3 4; do
    echo "variant $v layer $layer: $(RAFT_B200_PDL=0 RAFT_B200_MEGA=0 RAFT_B200_LIB=$PWD/tools/epi_exp/tc_$v.so timeout 120 python tools/timeline.py $layer 2>&1 | tail -n 3 | tr '\n' ' ')"
  done
done
