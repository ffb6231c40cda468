#!/bin/bash
# 8-GPU box: the driver's scaling invocation for the default config at N = 8 and 4, cfg5 at N = 8, reference arm at N = 8
OUT=gpurun_out
TAG=${1:-r02z}
mkdir -p $OUT
nvidia-smi -L | wc -l
run() {  # config nproc steps
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $2 --config $1 --steps $3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$1_n$2.json 2> $OUT/${TAG}_bench_$1_n$2.err
  python -c "
import json,sys
try:
    d=json.loads(open('$OUT/${TAG}_bench_$1_n$2.json').read().strip().splitlines()[-1]);print('$1 N=$2', 'value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']),d['clocks'])
except Exception as e: print('$1 N=$2 FAILED', e)"
  grep -v Warning $OUT/${TAG}_bench_$1_n$2.err | grep -i "error\|Traceback" | head -3
}
run cfg2 8 8
run cfg2 4 8
run cfg5 8 5
