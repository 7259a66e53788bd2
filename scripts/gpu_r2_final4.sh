#!/bin/bash
# verification of the final tree: full GPU suite, smoke, default bench, reference arm, launch list
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final4_pytest.log 2>&1
rc=$?
tail -n 4 gpurun_out/final4_pytest.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/final4_pytest.log | head; echo "tests failed rc=$rc"; exit 1; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -n 2
( time timeout 600 python bench.py > gpurun_out/final4_bench_default.json 2> gpurun_out/final4_bench_default.err ) 2> gpurun_out/final4_bench_time.txt
python - <<'PY'
import json
d = json.loads(open('gpurun_out/final4_bench_default.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel"], "fp64", d["roofline"]["fp64_secondary"]["frac"])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "seq", d["e2e"]["sequential"]["ms_per_step"])
print("hbm", d["roofline"]["hbm_bound_workload"]["frac"], d["roofline"]["hbm_bound_workload"]["kernel_ms"])
print("parity", d["parity_checks"], "clocks", d["clocks"], "launches", d["gpu_launches"])
for k, v in d["other_workloads"].items(): print(" ", k, v.get("ms_per_step", v.get("ms")), v.get("kernel"), v.get("trajectories_per_s"), v.get("trajectories_per_s_batch_32768"), v.get("parity_vs_oracle"))
print("cfg5", d["cfg5"]["ms_per_step"], d["cfg5"]["value"])
PY
tail -n 3 gpurun_out/final4_bench_time.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | tail -n 1 | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final4_launch_bench.log 2>&1
grep -c "k_stage\|k_rollout" gpurun_out/r02_launches_default_bench.csv
