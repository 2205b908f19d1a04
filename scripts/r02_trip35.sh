#!/bin/bash
# Round-2 trip 35: compute-sanitizer memcheck over small forwards of both variants and one paired update-block application.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 compute-sanitizer --tool memcheck --launch-timeout 120 python scripts/sanitize.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck (small forwards) exit $?"
grep -E "ok|ERROR SUMMARY" gpurun_out/r02_sanitizer_memcheck.log | tail -n 4
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 300 python scripts/sanitize_pair.py > gpurun_out/r02_sanitizer_memcheck_pair.log 2>&1; echo "memcheck (paired update block) exit $?"
grep -E "ok|ERROR SUMMARY|=========" gpurun_out/r02_sanitizer_memcheck_pair.log | head -n 12
