#!/bin/bash
# last verification of the round: full GPU suite, smoke, default bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final2_pytest.log 2>&1
rc=$?
tail -n 4 gpurun_out/final2_pytest.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/final2_pytest.log | head; echo "tests failed rc=$rc"; exit 1; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -n 2
( time timeout 900 python bench.py > gpurun_out/final2_bench_default.json 2> gpurun_out/final2_bench_default.err ) 2> gpurun_out/final2_bench_time.txt
python - <<'PY'
import json
d = json.loads(open('gpurun_out/final2_bench_default.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel"], "fp64", d["roofline"]["fp64_secondary"]["frac"])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "seq", d["e2e"]["sequential"]["ms_per_step"])
print("hbm", d["roofline"]["hbm_bound_workload"]["frac"], d["roofline"]["hbm_bound_workload"]["kernel_ms"])
print("parity", d["parity_checks"], "clocks", d["clocks"])
for k, v in d["other_workloads"].items(): print(" ", k, v.get("ms_per_step", v.get("ms")), v.get("kernel"))
print("cfg5", d["cfg5"]["ms_per_step"], d["cfg5"]["value"])
PY
tail -n 3 gpurun_out/final2_bench_time.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | tail -n 1 | cut -c1-400
B="python bench.py --no-cpu-baseline --no-e2e --no-others --steps 3 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_wide -s 3 -c 1 -f -o gpurun_out/r02_wide_kirk \
  $B > gpurun_out/final2_ncu.log 2>&1
ncu -i gpurun_out/r02_wide_kirk.ncu-rep --page raw --csv > gpurun_out/r02_wide_kirk_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02_wide_kirk_raw.csv > gpurun_out/r02_wide_kirk_summary.txt
cat gpurun_out/r02_wide_kirk_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final2_launch_bench.log 2>&1
grep -c k_stage gpurun_out/r02_launches_default_bench.csv
