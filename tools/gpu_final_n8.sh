#!/bin/bash
# final 8-GPU lines: the default weak-scaling workload and BASELINE configs[4] (2^34 pairs)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"
  grep '^{' gpurun_out/$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2),d['clocks'])
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d['exchange'].items() if k!='note'})
print(d.get('e2e'))"
}
run bench_n${N} --steps 5 --warmup 3
run bench_n${N}_2p34 --steps 3 --warmup 2 --pairs-per-gpu '1<<31' --no-e2e
