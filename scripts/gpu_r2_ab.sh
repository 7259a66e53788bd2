#!/bin/bash
# A/B of the strip kernel: this tree vs the revision measured in call E (worktree under _old/)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/ab_bench.log
W=attitude_x16_3x16000x4800x3
for rep in 1 2; do
  for tree in new old; do
    if [ $tree = old ]; then dir=_old; else dir=.; fi
    echo "== $tree 20 steps" >> gpurun_out/ab_bench.log
    (cd $dir && timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-others --workload $W) >> gpurun_out/ab_bench.log 2>&1
    echo "== $tree 300 steps" >> gpurun_out/ab_bench.log
    (cd $dir && timeout 200 python bench.py --steps 300 --warmup 50 --no-cpu-baseline --no-e2e --no-others --workload $W) >> gpurun_out/ab_bench.log 2>&1
  done
done
grep -E "== |ms_per_step" gpurun_out/ab_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"sm_mhz": ([0-9.a-z]+).*"kernel": "([a-z:]+)".*/  \1 ms  sm \2 \3/'
echo "--- stream x4 new / old"
for tree in new old; do
  if [ $tree = old ]; then dir=_old; else dir=.; fi
  (cd $dir && timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-others --workload pos_att_x4_120x120x80x60x9) 2>&1 | grep -o '"ms_per_step": [0-9.]*'
done
timeout 300 python -m pytest tests/test_gpu_idx_bytes.py tests/test_gpu_parity.py -m gpu -x -q -k "idx or strip or stream" 2>&1 | tail -n 2
