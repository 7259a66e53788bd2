#!/bin/bash
# round 2, GPU call C: stream kernel (4 full/empty barriers + trap watchdog): quick tests first, then ncu
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or pos_att" > gpurun_out/c_pytest_stream.log 2>&1
rc=$?
echo "stream tests exit $rc" >> gpurun_out/c_pytest_stream.log
tail -n 3 gpurun_out/c_pytest_stream.log
[ $rc -ne 0 ] && exit 1
timeout 300 python -m pytest tests/test_gpu_orbit.py tests/test_gpu_full_horizon.py -m gpu -x -q > gpurun_out/c_pytest_orbit.log 2>&1
echo "orbit/full-horizon tests exit $?" >> gpurun_out/c_pytest_orbit.log
tail -n 5 gpurun_out/c_pytest_orbit.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-others"
W4=pos_att_x4_120x120x80x60x9
: > gpurun_out/c_bench.log
for cfg in default 2,10,0,3 2,8,0,3; do
  if [ "$cfg" = default ]; then unset BELLMAN_STREAM; else export BELLMAN_STREAM=$cfg; fi
  echo "== x4 stream cfg $cfg" >> gpurun_out/c_bench.log
  timeout 120 $B --workload $W4 >> gpurun_out/c_bench.log 2>&1
done
unset BELLMAN_STREAM
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stage_stream -s 4 -c 1 -f -o gpurun_out/r02_stream5_posatt4 \
  $B --workload $W4 > gpurun_out/c_ncu.log 2>&1
ncu -i gpurun_out/r02_stream5_posatt4.ncu-rep --page raw --csv > gpurun_out/r02_stream5_posatt4_raw.csv 2>/dev/null
grep -E "== |ms_per_step" gpurun_out/c_bench.log | sed -E 's/.*"ms_per_step": ([0-9.]+).*"kernel": "([a-z:]+)".*/  \1 ms \2/'
tail -n 3 gpurun_out/c_ncu.log | cut -c1-300
