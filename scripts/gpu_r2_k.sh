#!/bin/bash
# k_stage_wide: chunk-size sweep on cfg 4 (one CTA per SM may use up to 220 KB for the two ring slots)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py tests/test_gpu_idx_bytes.py -m gpu -x -q \
  -k "window or kirk or group or narrow" > gpurun_out/k_pytest.log 2>&1
rc=$?
tail -n 5 gpurun_out/k_pytest.log
if [ $rc -ne 0 ]; then echo "tests failed rc=$rc"; exit 1; fi
B="python bench.py --no-cpu-baseline --no-e2e --no-others --steps 10 --warmup 3"
: > gpurun_out/k_bench.log
for cc in default 16 20; do
  echo "== kirk wide cc=$cc" >> gpurun_out/k_bench.log
  if [ $cc = default ]; then timeout 300 $B >> gpurun_out/k_bench.log 2>&1
  else BELLMAN_WIN_CC=$cc timeout 300 $B >> gpurun_out/k_bench.log 2>&1; fi
done
for cc in 26 16; do
echo "== kirk wide cc=$cc CTA barrier" >> gpurun_out/k_bench.log
BELLMAN_WIN_CC=$cc BELLMAN_WIDE_BARRIER=1 timeout 300 $B >> gpurun_out/k_bench.log 2>&1
done
grep -E "== |ms_per_step" gpurun_out/k_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"sm_mhz": ([0-9.a-z]+).*"kernel": "([a-z:]+)".*/  \1 ms  sm \2 \3/'
