#!/bin/bash
# round 2, GPU call A: validate + measure the streaming pos-att kernel (1 GPU)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile or stream or pos_att" > gpurun_out/a_pytest_stream.log 2>&1
echo "stream tests exit $?" >> gpurun_out/a_pytest_stream.log
timeout 900 python -m pytest tests/test_gpu_full_horizon.py -m gpu -x -q -s > gpurun_out/a_pytest_full.log 2>&1
echo "full-horizon tests exit $?" >> gpurun_out/a_pytest_full.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-others"
W4=pos_att_x4_120x120x80x60x9
W8=pos_att_x8_1ch_240x240x160x120x9
for cfg in default 2,4,0,3 2,6,0,3 2,10,0,3 1,4,0,3 4,4,0,3 2,4,0,4 2,3,0,3 2,2,0,3; do
  if [ "$cfg" = default ]; then unset BELLMAN_STREAM; else export BELLMAN_STREAM=$cfg; fi
  echo "== stream cfg $cfg" >> gpurun_out/a_bench.log
  BELLMAN_TILE_DEBUG=1 timeout 300 $B --workload $W4 >> gpurun_out/a_bench.log 2>&1
done
unset BELLMAN_STREAM
echo "== tile (old kernel)" >> gpurun_out/a_bench.log
BELLMAN_NO_STREAM=1 timeout 300 $B --workload $W4 >> gpurun_out/a_bench.log 2>&1
for cfg in default 2,2,0,3 2,4,0,3 2,6,0,3; do
  if [ "$cfg" = default ]; then unset BELLMAN_STREAM; else export BELLMAN_STREAM=$cfg; fi
  echo "== x8 stream cfg $cfg" >> gpurun_out/a_bench.log
  BELLMAN_TILE_DEBUG=1 timeout 300 $B --workload $W8 >> gpurun_out/a_bench.log 2>&1
done
unset BELLMAN_STREAM
echo "== x8 tile (old kernel)" >> gpurun_out/a_bench.log
BELLMAN_NO_STREAM=1 timeout 300 $B --workload $W8 >> gpurun_out/a_bench.log 2>&1
# ncu: full capture of one stream-kernel launch on pos-att x4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_stream -s 4 -c 1 -f -o gpurun_out/r02_stream_posatt4 \
  $B --workload $W4 > gpurun_out/a_ncu.log 2>&1
ncu -i gpurun_out/r02_stream_posatt4.ncu-rep --page raw --csv > gpurun_out/r02_stream_posatt4_raw.csv 2>/dev/null
# the rest of the GPU suite
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest_all.log 2>&1
echo "all gpu tests exit $?" >> gpurun_out/a_pytest_all.log
tail -3 gpurun_out/a_pytest_stream.log gpurun_out/a_pytest_full.log gpurun_out/a_pytest_all.log
grep -E "== |ms_per_step" gpurun_out/a_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"kernel": "([a-z:]+)".*/  \1 ms \2/' 
