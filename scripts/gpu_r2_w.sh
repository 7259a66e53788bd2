#!/bin/bash
# persistent sweep kernel: 1024-thread CTAs (a quarter of the grid-barrier arrivals) vs 256-thread CTAs
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -x -q -k "position or persistent or golden or full_horizon or facade" > gpurun_out/w_pytest.log 2>&1
echo "pytest exit $?"; tail -n 2 gpurun_out/w_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-others --steps 2000 --warmup 3 --workload position_3x201x201x3"
for rep in 1 2; do
  echo "== 1024"; timeout 200 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"kernel": "[a-z:]*"' | head -2 | tr '\n' ' '; echo
  echo "== 256"; BELLMAN_PERSIST_BLOCK256=1 timeout 200 $B 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"kernel": "[a-z:]*"' | head -2 | tr '\n' ' '; echo
done
