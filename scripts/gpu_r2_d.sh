#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream" > gpurun_out/d_pytest_stream.log 2>&1
rc=$?
echo "stream tests exit $rc" >> gpurun_out/d_pytest_stream.log
tail -n 3 gpurun_out/d_pytest_stream.log
[ $rc -ne 0 ] && exit 1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-others"
W4=pos_att_x4_120x120x80x60x9
W8=pos_att_x8_1ch_240x240x160x120x9
: > gpurun_out/d_bench.log
for cfg in default 2,4,0,3,2 2,4,0,3,3 2,6,0,3,3 2,8,0,3,4 2,10,0,3,2 2,10,0,3,0; do
  if [ "$cfg" = default ]; then unset BELLMAN_STREAM; else export BELLMAN_STREAM=$cfg; fi
  echo "== x4 stream cfg $cfg" >> gpurun_out/d_bench.log
  BELLMAN_TILE_DEBUG=1 timeout 120 $B --workload $W4 >> gpurun_out/d_bench.log 2>&1
done
for cfg in default 2,6,0,3,3 2,6,0,3,2 2,4,0,3,2; do
  if [ "$cfg" = default ]; then unset BELLMAN_STREAM; else export BELLMAN_STREAM=$cfg; fi
  echo "== x8 stream cfg $cfg" >> gpurun_out/d_bench.log
  BELLMAN_TILE_DEBUG=1 timeout 120 $B --workload $W8 >> gpurun_out/d_bench.log 2>&1
done
unset BELLMAN_STREAM
grep -E "== |ms_per_step|bellman stream" gpurun_out/d_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"kernel": "([a-z:]+)".*/  \1 ms \2/; s/bellman stream: (T = [0-9 ]+),.*smem = ([0-9]+ KB), ([0-9]+ threads \(NP = [0-9]+\)).*/  \1 \2 \3/'
