#!/bin/bash
mkdir -p gpurun_out
{
for v in prev default prev default; do
  if [ "$v" = default ]; then unset MSB64_B200_LIB; else export MSB64_B200_LIB=$PWD/inplacemsdradixsort_b200/lib/variants/libmsb64_$v.so; fi
  echo "== $v uniform 2^30"; timeout 300 python tools/dev_bench.py '1<<30' 0 0 2>&1 | grep -E "^[23] |bad|levels"
done
unset MSB64_B200_LIB
echo "== default dup1e6"; timeout 300 python tools/dev_bench.py '1<<30' 2 1000000 2>&1 | grep -E "^[3] |bad"
echo "== default sorted"; timeout 300 python tools/dev_bench.py '1<<30' 3 1 2>&1 | grep -E "^[3] |bad"
echo "== range 2^25 w56"; timeout 300 python tools/dev_range.py '1<<25' 56 32 2>&1 | tail -5
} > gpurun_out/ab8.log 2>&1
cat gpurun_out/ab8.log | cut -c1-230
timeout 600 ncu --set full --clock-control none --import-source on -k regex:local_sort_packed -c 1 \
  -o gpurun_out/r02_local -f python tools/dev_bench.py '1<<30' 0 0 > gpurun_out/ncu_local.log 2>&1; echo "ncu local rc=$?"
