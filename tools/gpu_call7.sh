#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_parity.py -m gpu -x -q -k "sharing or single_gpu or pipelined_sizes or dropin or over_shards" > gpurun_out/pytest_shard.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_shard.log
