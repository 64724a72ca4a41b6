#!/bin/bash
mkdir -p gpurun_out
{
for v in cap2k_4 cap2k_5; do
  export MSB64_B200_LIB=$PWD/inplacemsdradixsort_b200/lib/variants/libmsb64_$v.so
  echo "== $v uniform 2^30"; timeout 300 python tools/dev_bench.py '1<<30' 0 0 2>&1 | tail -6
  echo "== $v uniform 2^28"; timeout 300 python tools/dev_bench.py '1<<28' 0 0 2>&1 | tail -5
  echo "== $v dup1e6 2^30"; timeout 300 python tools/dev_bench.py '1<<30' 2 1000000 2>&1 | tail -5
  echo "== $v range 2^26 w57"; timeout 300 python tools/dev_range.py '1<<26' 57 16 2>&1 | tail -6
done
} > gpurun_out/ab4.log 2>&1
cat gpurun_out/ab4.log | cut -c1-260
