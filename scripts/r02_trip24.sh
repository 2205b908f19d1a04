#!/bin/bash
# Round-2 trip 24: per-chunk timelines of the narrow update-block layers (per-layer kernel: same mainloop as the mega kernel).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 RAFT_B200_MEGA=0
for layer in 9 3 5 2; do
  timeout 120 python tools/timeline.py $layer > gpurun_out/r02_timeline_layer$layer.log 2>&1
  echo "=== layer $layer"; head -n 22 gpurun_out/r02_timeline_layer$layer.log | tail -n 21; tail -n 16 gpurun_out/r02_timeline_layer$layer.log
done
