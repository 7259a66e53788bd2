#!/bin/bash
# k_stage_wide: controls unrolled per loop iteration (library variants built with -DWIDE_UNROLL=1 / 4)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
P=optimal-control-dynamic-programming_b200
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window_kernel" > gpurun_out/q_pytest.log 2>&1
rc=$?
tail -n 3 gpurun_out/q_pytest.log
if [ $rc -ne 0 ]; then echo "tests failed rc=$rc"; exit 1; fi
B="python bench.py --no-cpu-baseline --no-e2e --no-others --steps 10 --warmup 3"
: > gpurun_out/q_bench.log
cp $P/libbellman.so /tmp/libbellman_default.so
for rep in 1 2; do
  for v in default u1 u4; do
    if [ $v = default ]; then cp /tmp/libbellman_default.so $P/libbellman.so; else cp $P/libbellman_$v.so.variant $P/libbellman.so; fi
    echo "== kirk wide unroll $v" >> gpurun_out/q_bench.log
    timeout 300 $B >> gpurun_out/q_bench.log 2>&1
  done
done
cp /tmp/libbellman_default.so $P/libbellman.so
grep -E "== |ms_per_step" gpurun_out/q_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"sm_mhz": ([0-9.a-z]+).*"kernel": "([a-z:]+)".*/  \1 ms  sm \2 \3/'
