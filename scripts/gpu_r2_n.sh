#!/bin/bash
# N-GPU call: sharded bellman_stage_host check + one sharded sweep check, then the bench at N
set -u
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
: > gpurun_out/n${N}_check.log
port=29500
for kind in kirk_host kirk_odd; do
  port=$((port+1))
  echo "== $kind p2p" >> gpurun_out/n${N}_check.log
  timeout 300 $TR --master-port $port scripts/multi_gpu_check.py $kind >> gpurun_out/n${N}_check.log 2>&1
  echo "exit $?" >> gpurun_out/n${N}_check.log
done
grep -E "== |MULTI_GPU_CHECK|exit|MISMATCH|stage_host" gpurun_out/n${N}_check.log | cut -c1-160
if grep -q "MISMATCH\|FAIL" gpurun_out/n${N}_check.log; then echo "check failed"; exit 1; fi
export BELLMAN_BENCH_VERBOSE=1
timeout 900 $TR --master-port 29590 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo "bench exit $?"
grep -E "^rank" gpurun_out/n${N}_bench.err | tail -20
python - <<PY
import json
for l in open("gpurun_out/n${N}_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print({k: d[k] for k in ("value", "ms_per_step", "exchange_ms_per_step", "sharded_parity", "parity_checks", "host_cpus_bound_per_rank")}, d.get("run", d["config"]).get("slab_cuts"))
        print("  cfg5", d.get("cfg5"))
        print("  e2e", d.get("e2e"))
PY
