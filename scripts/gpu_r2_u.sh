#!/bin/bash
# after the adaptive slab count: the host-stage tests, then the default bench (the line the driver will see)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/u_pytest.log 2>&1
echo "pytest exit $?"; tail -n 2 gpurun_out/u_pytest.log
timeout 900 python bench.py > gpurun_out/u_bench_default.json 2> gpurun_out/u_bench_default.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/u_bench_default.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel"], "fp64", d["roofline"]["fp64_secondary"]["frac"])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "seq", d["e2e"]["sequential"]["ms_per_step"])
print("hbm", d["roofline"]["hbm_bound_workload"]["frac"], d["roofline"]["hbm_bound_workload"]["kernel_ms"])
print("parity", d["parity_checks"], "clocks", d["clocks"])
PY
