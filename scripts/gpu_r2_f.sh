#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_idx_bytes.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/f_pytest_idx.log 2>&1; tail -n 5 gpurun_out/f_pytest_idx.log
timeout 120 python scripts/sanitize_small.py > gpurun_out/f_sanitize_plain.log 2>&1; echo "plain exit $?"; tail -n 3 gpurun_out/f_sanitize_plain.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r02_racecheck_small.log 2>&1
echo "racecheck exit $?"; tail -n 12 gpurun_out/r02_racecheck_small.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r02_memcheck_small.log 2>&1
echo "memcheck exit $?"; tail -n 6 gpurun_out/r02_memcheck_small.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-others"
for W in pos_att_x4_120x120x80x60x9 pos_att_x8_1ch_240x240x160x120x9; do
  tag=$(echo $W | cut -d_ -f3)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stage_stream -s 4 -c 1 -f -o gpurun_out/r02_stream_posatt_$tag \
    $B --workload $W > gpurun_out/f_ncu_$tag.log 2>&1
  ncu -i gpurun_out/r02_stream_posatt_$tag.ncu-rep --page raw --csv > gpurun_out/r02_stream_posatt_${tag}_raw.csv 2>/dev/null
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_launch_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r02_launches_default_bench.csv
