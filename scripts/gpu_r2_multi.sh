#!/bin/bash
# round 2, multi-GPU call (N = $1): sharded parity on small grids (both halo modes), then the bench at N
set -u
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
: > gpurun_out/m${N}_check.log
port=29600
for kind in kirk kirk_odd attitude pos_att; do
  for mode in p2p nccl; do
    port=$((port+1))
    echo "== $kind $mode" >> gpurun_out/m${N}_check.log
    if [ $mode = nccl ]; then export BELLMAN_NO_P2P=1; else unset BELLMAN_NO_P2P; fi
    timeout 300 $TR --master-port $port scripts/multi_gpu_check.py $kind >> gpurun_out/m${N}_check.log 2>&1
    echo "exit $?" >> gpurun_out/m${N}_check.log
  done
done
unset BELLMAN_NO_P2P
grep -E "== |MULTI_GPU_CHECK|exit|MISMATCH" gpurun_out/m${N}_check.log
export BELLMAN_BENCH_VERBOSE=1
timeout 900 $TR --master-port 29700 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
echo "bench exit $?"
tail -c 3000 gpurun_out/m${N}_bench.json
grep -E "rank|Error|error" gpurun_out/m${N}_bench.err | tail -20
