#!/bin/bash
# extra sanitizer tools on the small cases + the GPU tests added late in the round
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/o_pytest.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/o_pytest.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 10 python scripts/sanitize_small.py > gpurun_out/r02_synccheck_small.log 2>&1
echo "synccheck exit $?"; tail -n 4 gpurun_out/r02_synccheck_small.log
timeout 900 compute-sanitizer --tool initcheck --print-limit 10 python scripts/sanitize_small.py > gpurun_out/r02_initcheck_small.log 2>&1
echo "initcheck exit $?"; tail -n 6 gpurun_out/r02_initcheck_small.log
grep -c "Uninitialized" gpurun_out/r02_initcheck_small.log
