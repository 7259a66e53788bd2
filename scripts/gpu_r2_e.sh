#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q > gpurun_out/e_pytest_group.log 2>&1
echo "group tests exit $?" >> gpurun_out/e_pytest_group.log
tail -n 15 gpurun_out/e_pytest_group.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/e_pytest_all.log 2>&1
echo "all gpu tests exit $?" >> gpurun_out/e_pytest_all.log
tail -n 6 gpurun_out/e_pytest_all.log
( time timeout 900 python bench.py > gpurun_out/e_bench_default.json 2> gpurun_out/e_bench_default.err ) 2> gpurun_out/e_bench_time.txt
echo "bench exit $?"; cat gpurun_out/e_bench_time.txt
tail -c 2500 gpurun_out/e_bench_default.json; tail -5 gpurun_out/e_bench_default.err
