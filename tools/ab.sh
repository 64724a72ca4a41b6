#!/bin/bash
# tools/ab.sh N KIND PARAM SCHED VARIANT...   -- dev_bench of several library variants ("default" = in-tree lib)
n=$1; kind=$2; param=$3; sched=$4; shift 4
for v in "$@"; do
  if [ "$v" = default ]; then unset MSB64_B200_LIB; else export MSB64_B200_LIB=$PWD/inplacemsdradixsort_b200/lib/variants/libmsb64_$v.so; fi
  echo "== $v n=$n kind=$kind sched=$sched"
  python tools/dev_bench.py "$n" "$kind" "$param" $sched 2>&1 | grep -E "^3 |bad|best"
done
