#!/bin/bash
mkdir -p gpurun_out
{
echo "== uniform 2^30"; timeout 300 python tools/dev_bench.py '1<<30' 0 0 2>&1 | tail -7
echo "== dup1e6 2^30"; timeout 300 python tools/dev_bench.py '1<<30' 2 1000000 2>&1 | tail -6
echo "== sorted 2^30"; timeout 300 python tools/dev_bench.py '1<<30' 3 1 2>&1 | tail -6
echo "== low24 2^30"; timeout 300 python tools/dev_bench.py '1<<30' 1 16777215 2>&1 | tail -6
echo "== range 2^26 w57"; timeout 300 python tools/dev_range.py '1<<26' 57 16 2>&1 | tail -7
echo "== range 2^25 w56"; timeout 300 python tools/dev_range.py '1<<25' 56 32 2>&1 | tail -7
} > gpurun_out/ab5.log 2>&1
cat gpurun_out/ab5.log | cut -c1-250
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
