#!/bin/bash
# bellman_stage_host: parity test, then the default workload's e2e (pipelined vs sequential)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_host or window_kernel" > gpurun_out/l_pytest.log 2>&1
rc=$?
tail -n 5 gpurun_out/l_pytest.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/l_pytest.log | head -20; echo "tests failed rc=$rc"; exit 1; fi
timeout 600 python bench.py --no-cpu-baseline --no-others --steps 10 --warmup 3 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/l_bench.json').read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "kernel", d["roofline"]["kernel"])
print("e2e", json.dumps(d["e2e"]))
print("parity", d["parity_checks"])
PY
tail -n 3 gpurun_out/l_bench.err
