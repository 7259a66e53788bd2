#!/bin/bash
# round 2, GPU call B: stream kernel v3 (mbarrier pipeline) — tests, geometry sweep, ncu
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_horizon.py -m gpu -x -q -k "tile or stream or pos_att" > gpurun_out/b_pytest_stream.log 2>&1
echo "stream tests exit $?" >> gpurun_out/b_pytest_stream.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-others"
W4=pos_att_x4_120x120x80x60x9
W8=pos_att_x8_1ch_240x240x160x120x9
: > gpurun_out/b_bench.log
for cfg in default 2,4,0,2 2,6,0,3 2,8,0,3 2,10,0,3 1,4,0,3 4,4,0,3 2,3,0,3 2,2,0,3; do
  if [ "$cfg" = default ]; then unset BELLMAN_STREAM; else export BELLMAN_STREAM=$cfg; fi
  echo "== x4 stream cfg $cfg" >> gpurun_out/b_bench.log
  BELLMAN_TILE_DEBUG=1 timeout 300 $B --workload $W4 >> gpurun_out/b_bench.log 2>&1
done
for cfg in default 2,2,0,3 2,3,0,3 2,4,0,3 2,6,0,2; do
  if [ "$cfg" = default ]; then unset BELLMAN_STREAM; else export BELLMAN_STREAM=$cfg; fi
  echo "== x8 stream cfg $cfg" >> gpurun_out/b_bench.log
  BELLMAN_TILE_DEBUG=1 timeout 300 $B --workload $W8 >> gpurun_out/b_bench.log 2>&1
done
unset BELLMAN_STREAM
echo "== ref-size stream" >> gpurun_out/b_bench.log
timeout 300 $B --workload pos_att_ref_30x30x20x15x9 --steps 50 >> gpurun_out/b_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_stream -s 4 -c 1 -f -o gpurun_out/r02_stream3_posatt4 \
  $B --workload $W4 > gpurun_out/b_ncu.log 2>&1
ncu -i gpurun_out/r02_stream3_posatt4.ncu-rep --page raw --csv > gpurun_out/r02_stream3_posatt4_raw.csv 2>/dev/null
tail -n 3 gpurun_out/b_pytest_stream.log
grep -E "== |ms_per_step|bellman stream" gpurun_out/b_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"kernel": "([a-z:]+)".*/  \1 ms \2/'
