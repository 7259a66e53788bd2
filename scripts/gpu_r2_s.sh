#!/bin/bash
# sanity after reverting the strip experiment: full GPU suite + the two headline numbers
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1
echo "pytest exit $?"; tail -n 2 gpurun_out/s_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-others --warmup 3"
for w in attitude_x16_3x16000x4800x3 kirk_scaled_8192x8192x512; do
  timeout 300 $B --steps 20 --workload $w 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"kernel": "[a-z:]*"' | head -2 | tr '\n' ' '; echo
done
