#!/bin/bash
# 8-GPU round: the direct-store fraction of the route pass (eighths), then the two bench lines
N=${1:-8}
mkdir -p gpurun_out
run() { # name, extra args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"
  grep '^{' gpurun_out/$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2),d['clocks'])
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d['exchange'].items() if k!='note'})
print(d.get('e2e'))"
}
best=2; bestv=0
for e in 2 3 4; do
  export MSB64_SHARD_DIRECT_EIGHTHS=$e
  run bench_n${N}_d$e --steps 3 --warmup 2 --no-e2e
  v=$(grep '^{' gpurun_out/bench_n${N}_d$e.json | python -c "import json,sys; print(json.loads(sys.stdin.readline())['value'])" 2>/dev/null || echo 0)
  if python -c "import sys; sys.exit(0 if float('$v') > float('$bestv') else 1)"; then best=$e; bestv=$v; fi
done
echo "best eighths=$best ($bestv Gpairs/s)" | tee gpurun_out/n${N}_best.txt
export MSB64_SHARD_DIRECT_EIGHTHS=$best
run bench_n${N} --steps 5 --warmup 3
run bench_n${N}_2p34 --steps 3 --warmup 2 --pairs-per-gpu '1<<31' --no-e2e
