#!/bin/bash
# sharded bellman_stage_host: slab count at N ranks
set -u
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29300
for n in 8 4 16 12; do
  port=$((port+1))
  echo "== N=$N slabs $n"
  BELLMAN_HOST_SLABS=$n timeout 300 $TR --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --no-others --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  kernel %.2f ms  e2e %.2f ms  sequential %.2f ms  %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['sequential']['ms_per_step'], d['parity_checks']))
"
done
