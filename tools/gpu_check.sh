#!/bin/bash
# tools/gpu_check.sh -- the standard single-GPU round on a gpurun box: GPU tests, a developer
# timing of the sort, the smoke entry point.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/dev_bench.py '1<<30' 0 0 2>&1 | grep -E "^[23] |bad|levels" | cut -c1-220
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
