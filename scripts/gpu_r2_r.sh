#!/bin/bash
# (historical: the multi-tile variant measured here lost and was reverted; BELLMAN_STRIP_K no longer exists)
# strip kernel: consecutive dimension-1 tiles per CTA (BELLMAN_STRIP_K)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_horizon.py tests/test_gpu_group.py tests/test_gpu_idx_bytes.py -m gpu -x -q -k "strip or attitude or narrow" > gpurun_out/r_pytest.log 2>&1
rc=$?
tail -n 3 gpurun_out/r_pytest.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/r_pytest.log | head; echo "tests failed rc=$rc"; exit 1; fi
B="python bench.py --no-cpu-baseline --no-e2e --no-others --steps 20 --warmup 3 --workload"
: > gpurun_out/r_bench.log
run() { echo "== $*" >> gpurun_out/r_bench.log; ( env "${@:2}" timeout 200 $B $1 ) >> gpurun_out/r_bench.log 2>&1; }
for rep in 1 2; do
  W=attitude_x16_3x16000x4800x3
  run $W X=1
  for k in 2 3 4 6 8; do run $W BELLMAN_STRIP_K=$k; done
  run $W BELLMAN_STRIP_K=4 BELLMAN_STRIP_PF=1184
  run $W BELLMAN_STRIP_K=4 BELLMAN_STRIP_PF=296
  run $W BELLMAN_STRIP_K=4 BELLMAN_STRIP_R=8
done
W=attitude_x4_3x4000x1200x3
for k in 1 2 4; do run $W BELLMAN_STRIP_K=$k; done
grep -E "== |ms_per_step" gpurun_out/r_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"sm_mhz": ([0-9.a-z]+).*"kernel": "([a-z:]+)".*/  \1 ms  sm \2 \3/'
