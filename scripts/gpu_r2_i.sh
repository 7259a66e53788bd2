#!/bin/bash
# k_stage_wide (512-thread, 4 states per thread, constant-bank control tables) vs k_stage_window on cfg 4;
# strip kernel with distance-2 table prefetch (BELLMAN_WIN_OCC=3 variant) vs default
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py tests/test_gpu_idx_bytes.py -m gpu -x -q \
  -k "window or kirk or group or narrow or strip" > gpurun_out/i_pytest.log 2>&1
rc=$?
tail -n 5 gpurun_out/i_pytest.log
if [ $rc -ne 0 ]; then echo "tests failed rc=$rc"; exit 1; fi
B="python bench.py --no-cpu-baseline --no-e2e --no-others"
: > gpurun_out/i_bench.log
run() {  # label, env..., -- args
  local label=$1; shift
  echo "== $label" >> gpurun_out/i_bench.log
  ( env "$@" timeout 300 $B $ARGS ) >> gpurun_out/i_bench.log 2>&1
}
ARGS="--steps 10 --warmup 3"
for rep in 1 2; do
  run "kirk wide ns2" X=1
  run "kirk wide ns3" BELLMAN_WIDE_NS=3
  run "kirk ring" BELLMAN_NO_WIDE=1
done
ARGS="--steps 20 --warmup 3 --workload attitude_x16_3x16000x4800x3"
for rep in 1 2 3; do
  run "strip default" X=1
  run "strip occ5 pf2" BELLMAN_WIN_OCC=3
done
grep -E "== |ms_per_step" gpurun_out/i_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"sm_mhz": ([0-9.a-z]+).*"kernel": "([a-z:]+)".*/  \1 ms  sm \2 \3/'
