#!/bin/bash
# (historical: the BELLMAN_WIN_R4 variant measured here lost and was removed from the library afterwards)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-others"
: > gpurun_out/g_bench.log
run() { echo "== $*" >> gpurun_out/g_bench.log; env "$@" timeout 200 $B >> gpurun_out/g_bench.log 2>&1; }
run X=0
run BELLMAN_WIN_R4=1 BELLMAN_WIN_OCC=3
run BELLMAN_WIN_R4=1 BELLMAN_WIN_OCC=4
run BELLMAN_WIN_R4=1 BELLMAN_WIN_OCC=3 BELLMAN_WIN_CC=6
run BELLMAN_WIN_R4=1 BELLMAN_WIN_OCC=4 BELLMAN_WIN_CC=4
run BELLMAN_WIN_R4=1 BELLMAN_WIN_OCC=3 BELLMAN_WIN_CC=12
run BELLMAN_WIN_CC=6
run BELLMAN_WIN_CC=12
run BELLMAN_WIN_CC=16
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window or kirk" > gpurun_out/g_pytest.log 2>&1; tail -n 2 gpurun_out/g_pytest.log
BELLMAN_WIN_R4=1 BELLMAN_WIN_OCC=3 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window or kirk" > gpurun_out/g_pytest_r4.log 2>&1; tail -n 2 gpurun_out/g_pytest_r4.log
grep -E "== |ms_per_step" gpurun_out/g_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"kernel": "([a-z:]+)".*/  \1 ms \2/'
