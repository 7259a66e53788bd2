#!/bin/bash
# ncu --set full of one k_stage_wide launch on cfg 4 (Kirk 8192 x 8192 x 512), raw + source pages
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu-baseline --no-e2e --no-others --steps 3 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_wide -s 3 -c 1 -f -o gpurun_out/r02_wide_kirk \
  $B > gpurun_out/j_ncu.log 2>&1
tail -n 3 gpurun_out/j_ncu.log
ncu -i gpurun_out/r02_wide_kirk.ncu-rep --page raw --csv > gpurun_out/r02_wide_kirk_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_wide_kirk.ncu-rep --page source --csv > gpurun_out/r02_wide_kirk_source.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02_wide_kirk_raw.csv
