#!/bin/bash
# k_stage_wide: floor(g) as a double from the conversion pipe (I2F) for 0 / 1 / 2 dimensions
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window_kernel" > gpurun_out/p_pytest.log 2>&1
rc=$?
tail -n 3 gpurun_out/p_pytest.log
if [ $rc -ne 0 ]; then echo "tests failed rc=$rc"; exit 1; fi
B="python bench.py --no-cpu-baseline --no-e2e --no-others --steps 10 --warmup 3"
: > gpurun_out/p_bench.log
for rep in 1 2; do
  for xu in 0 1 2; do
    echo "== kirk wide xu=$xu" >> gpurun_out/p_bench.log
    BELLMAN_WIDE_XU=$xu timeout 300 $B >> gpurun_out/p_bench.log 2>&1
  done
done
grep -E "== |ms_per_step" gpurun_out/p_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"sm_mhz": ([0-9.a-z]+).*"kernel": "([a-z:]+)".*/  \1 ms  sm \2 \3/'
