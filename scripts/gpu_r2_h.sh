#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --steps 1500 --warmup 500 --no-cpu-baseline --no-e2e --no-others"
: > gpurun_out/h_bench.log
for rep in 1 2; do
for W in attitude_x16_3x16000x4800x3; do
  echo "== $W int32" >> gpurun_out/h_bench.log; timeout 200 $B --workload $W >> gpurun_out/h_bench.log 2>&1
  echo "== $W uint8" >> gpurun_out/h_bench.log; timeout 200 $B --workload $W --idx-bytes 1 >> gpurun_out/h_bench.log 2>&1
  echo "== $W int32 OCC3" >> gpurun_out/h_bench.log; BELLMAN_WIN_OCC=3 timeout 200 $B --workload $W >> gpurun_out/h_bench.log 2>&1
done
done
grep -E "== |ms_per_step" gpurun_out/h_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"frac": ([0-9.]+), "traffic.*"kernel": "([a-z:]+)".*/  \1 ms frac \2 \3/'
grep -o '"clocks": {[^}]*}' gpurun_out/h_bench.log | head -3
