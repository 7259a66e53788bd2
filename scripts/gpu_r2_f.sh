#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_idx_bytes.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/f_pytest_idx.log 2>&1; tail -n 5 gpurun_out/f_pytest_idx.log
timeout 120 python scripts/sanitize_small.py > gpurun_out/f_sanitize_plain.log 2>&1; echo "plain exit $?"; tail -n 3 gpurun_out/f_sanitize_plain.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python scripts/sanitize_small.py > gpurun_out/r02_racecheck_small.log 2>&1
echo "racecheck exit $?"; tail -n 12 gpurun_out/r02_racecheck_small.log
BELLMAN_STREAM_DEBUG_SYNC=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python scripts/sanitize_small.py > gpurun_out/r02_racecheck_small_debugsync.log 2>&1
echo "racecheck (stream kernel with a CTA barrier per iteration) exit $?"; tail -n 12 gpurun_out/r02_racecheck_small_debugsync.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r02_memcheck_small.log 2>&1
echo "memcheck exit $?"; tail -n 6 gpurun_out/r02_memcheck_small.log
