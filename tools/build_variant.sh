#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG ...]  -- a tuning variant of the library for A/B timing
# (tools/dev_bench.py with MSB64_B200_LIB=inplacemsdradixsort_b200/lib/variants/libmsb64_NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p inplacemsdradixsort_b200/lib/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -cudart static "$@" \
  -o inplacemsdradixsort_b200/lib/variants/libmsb64_$name.so inplacemsdradixsort_b200/csrc/msb64_b200.cu
echo built $name
