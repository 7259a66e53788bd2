#!/bin/bash
# bellman_stage_host: number of slabs (copies overlap the kernel slab by slab)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu-baseline --no-others --steps 5 --warmup 3"
for n in 8 6 12 16 24; do
  echo "== slabs $n"
  BELLMAN_HOST_SLABS=$n timeout 300 $B 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  kernel %.2f ms  e2e %.2f ms  sequential %.2f ms  %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['sequential']['ms_per_step'], d['parity_checks']))
"
done
