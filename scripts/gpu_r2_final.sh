#!/bin/bash
# Round-2 verification of the tree as committed: full GPU test suite, smoke, default bench, ncu capture of
# the headline kernel (k_stage_wide), launch list of the default bench, sanitizer runs on the small cases.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1
rc=$?
tail -n 4 gpurun_out/final_pytest.log
if [ $rc -ne 0 ]; then echo "tests failed rc=$rc"; exit 1; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -n 2
( time timeout 900 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err ) 2> gpurun_out/final_bench_time.txt
tail -c 600 gpurun_out/final_bench_default.json; echo; tail -n 3 gpurun_out/final_bench_time.txt
B="python bench.py --no-cpu-baseline --no-e2e --no-others --steps 3 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_wide -s 3 -c 1 -f -o gpurun_out/r02_wide_kirk \
  $B > gpurun_out/final_ncu.log 2>&1
ncu -i gpurun_out/r02_wide_kirk.ncu-rep --page raw --csv > gpurun_out/r02_wide_kirk_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02_wide_kirk_raw.csv > gpurun_out/r02_wide_kirk_summary.txt
cat gpurun_out/r02_wide_kirk_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_launch_bench.log 2>&1
grep -c k_stage gpurun_out/r02_launches_default_bench.csv
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r02_memcheck_small.log 2>&1
echo "memcheck exit $?"; tail -n 4 gpurun_out/r02_memcheck_small.log
BELLMAN_STREAM_DEBUG_SYNC=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python scripts/sanitize_small.py > gpurun_out/r02_racecheck_small_debugsync.log 2>&1
echo "racecheck (debug sync) exit $?"; tail -n 4 gpurun_out/r02_racecheck_small_debugsync.log
