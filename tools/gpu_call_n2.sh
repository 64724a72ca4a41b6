#!/bin/bash
# N-GPU round: the 2-rank NCCL parity tests, then bench.py under torchrun.  Usage: gpu_call_n2.sh N [exchange...]
N=${1:-2}; shift
EX=${@:-pipelined}
mkdir -p gpurun_out
[ "$SKIPTESTS" = 1 ] || timeout 600 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus or dropin or over_shards" > gpurun_out/pytest_n$N.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_n$N.log
for ex in $EX; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 --exchange $ex > gpurun_out/bench_n${N}_$ex.json 2> gpurun_out/bench_n${N}_$ex.err; echo "bench $ex rc=$?"
  grep '^{' gpurun_out/bench_n${N}_$ex.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2),'launches',d['gpu_launches'],d['clocks'])
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d['exchange'].items() if k!='note'})
print(d.get('e2e'))"
done
