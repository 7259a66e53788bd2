#!/bin/bash
# strip kernel: columns per warp (tile 32 x 4R) decide the shared memory per CTA, hence the CTAs per SM
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu-baseline --no-e2e --no-others --steps 20 --warmup 3 --workload attitude_x16_3x16000x4800x3"
: > gpurun_out/m_bench.log
run() { echo "== $*" >> gpurun_out/m_bench.log; ( env "$@" timeout 200 $B ) >> gpurun_out/m_bench.log 2>&1; }
for rep in 1 2; do
  run X=1
  for r in 8 10 12 14 20; do run BELLMAN_STRIP_R=$r; done
  run BELLMAN_STRIP_R=12 BELLMAN_STRIP_PF=740
  run BELLMAN_STRIP_R=12 BELLMAN_STRIP_PF=592
  run BELLMAN_STRIP_R=10 BELLMAN_STRIP_PF=888
  run BELLMAN_STRIP_R=14 BELLMAN_STRIP_PF=592
done
grep -E "== |ms_per_step" gpurun_out/m_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"sm_mhz": ([0-9.a-z]+).*"kernel": "([a-z:]+)".*/  \1 ms  sm \2 \3/'
