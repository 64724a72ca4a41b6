#!/bin/bash
# final single-GPU round: GPU tests, both bench arms, the ncu launch list and --set full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-workloads --no-e2e --no-cpu \
  > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'scatter_kernel|local_sort_packed|histogram_kernel' -c 6 \
  -o gpurun_out/r02_full -f python tools/dev_bench.py '1<<30' 0 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tail_kernel -c 1 \
  -o gpurun_out/r02_tail -f python tools/dev_bench.py '1<<28' 3 1 > gpurun_out/ncu_tail.log 2>&1; echo "ncu tail rc=$?"
ls -la gpurun_out/*.ncu-rep
