#!/bin/bash
# round 2, N-GPU call: two sharded checks, then the bench at N (balanced and equal slabs)
set -u
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
: > gpurun_out/m${N}_check.log
port=29800
for kind in kirk_odd pos_att; do
  port=$((port+1))
  echo "== $kind p2p" >> gpurun_out/m${N}_check.log
  timeout 300 $TR --master-port $port scripts/multi_gpu_check.py $kind >> gpurun_out/m${N}_check.log 2>&1
  echo "exit $?" >> gpurun_out/m${N}_check.log
done
grep -E "== |MULTI_GPU_CHECK|exit|MISMATCH" gpurun_out/m${N}_check.log
export BELLMAN_BENCH_VERBOSE=1
timeout 900 $TR --master-port 29900 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
echo "bench exit $?"
grep -E "^rank" gpurun_out/m${N}_bench.err | tail -20
timeout 600 $TR --master-port 29901 bench.py --gpus $N --steps 20 --warmup 3 --no-balance --no-others --no-e2e > gpurun_out/m${N}_bench_equal.json 2> gpurun_out/m${N}_bench_equal.err
echo "bench (equal slabs) exit $?"
grep -E "^rank" gpurun_out/m${N}_bench_equal.err | tail -20
python - <<PY
import json
for f in ("gpurun_out/m${N}_bench.json", "gpurun_out/m${N}_bench_equal.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, {k: d[k] for k in ("value", "ms_per_step", "exchange_ms_per_step", "sharded_parity")}, d.get("run", d["config"]).get("slab_cuts"))
            print("  cfg5", d.get("cfg5"))
            print("  e2e", d.get("e2e"))
PY
