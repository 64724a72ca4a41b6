#!/bin/bash
# tools/gpu_check.sh -- the standard single-GPU round on a gpurun box: GPU tests, the bench line,
# the ncu launch list of the same command.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-workloads --no-e2e --no-cpu \
  > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
