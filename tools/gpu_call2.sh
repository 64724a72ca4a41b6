#!/bin/bash
# A/B of library variants + schedule variants + one ncu --set full capture (single GPU)
mkdir -p gpurun_out
{
for v in base split; do
  export MSB64_B200_LIB=$PWD/inplacemsdradixsort_b200/lib/variants/libmsb64_$v.so
  echo "== $v uniform 2^30"; timeout 300 python tools/dev_bench.py '1<<30' 0 0 2>&1 | grep -E "^[23] |bad|best|levels"
  echo "== $v uniform 2^30 sched 7,7,6"; timeout 300 python tools/dev_bench.py '1<<30' 0 0 7,7,6,8,8,8,8,8,4 2>&1 | grep -E "^[23] |bad|best|levels"
  echo "== $v dup1e6 2^30"; timeout 300 python tools/dev_bench.py '1<<30' 2 1000000 2>&1 | grep -E "^[3] |bad|best|levels"
  echo "== $v sorted 2^30"; timeout 300 python tools/dev_bench.py '1<<30' 3 1 2>&1 | grep -E "^[3] |bad|best|levels"
done
export MSB64_B200_LIB=$PWD/inplacemsdradixsort_b200/lib/variants/libmsb64_split.so
for sc in 8,11,8,8,8,8,8,5 9,10,8,8,8,8,8,5 10,9,8,8,8,8,8,5 8,8,8,8,8,8,8,8 6,6,7,8,8,8,8,8,5; do
  echo "== split uniform 2^30 sched $sc"; timeout 300 python tools/dev_bench.py '1<<30' 0 0 $sc 2>&1 | grep -E "^[3] |bad|best|levels"
done
for v in cap2k_4 cap2k_5; do
  export MSB64_B200_LIB=$PWD/inplacemsdradixsort_b200/lib/variants/libmsb64_$v.so
  echo "== $v uniform 2^30 sched 7,7,6"; timeout 300 python tools/dev_bench.py '1<<30' 0 0 7,7,6,8,8,8,8,8,4 2>&1 | grep -E "^[23] |bad|best|levels"
done
} > gpurun_out/ab2.log 2>&1
export MSB64_B200_LIB=$PWD/inplacemsdradixsort_b200/lib/variants/libmsb64_split.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 1 -c 1 \
  -o gpurun_out/r02_scatter -f python tools/dev_bench.py '1<<28' 0 0 > gpurun_out/ncu_scatter.log 2>&1; echo "ncu scatter rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:local_sort_packed -c 1 \
  -o gpurun_out/r02_local -f python tools/dev_bench.py '1<<28' 0 0 > gpurun_out/ncu_local.log 2>&1; echo "ncu local rc=$?"
cat gpurun_out/ab2.log
