#!/bin/bash
N=4
mkdir -p gpurun_out
for e in 2 3; do
  export MSB64_SHARD_DIRECT_EIGHTHS=$e
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 2 --no-e2e > gpurun_out/bench_n4_d$e.json 2> gpurun_out/bench_n4_d$e.err; echo "d$e rc=$?"
  grep '^{' gpurun_out/bench_n4_d$e.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2))
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d['exchange'].items() if k!='note'})"
done
