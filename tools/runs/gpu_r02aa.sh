#!/bin/bash
# 4-GPU box: does limiting NCCL's CTA count leave more SMs to the backward GEMMs it overlaps with?
OUT=gpurun_out
TAG=${1:-r02aa}
mkdir -p $OUT
run() {  # label env
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus 4 --config cfg2 --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$1.json 2> $OUT/${TAG}_bench_$1.err
  python -c "
import json
try:
    d=json.loads(open('$OUT/${TAG}_bench_$1.json').read().strip().splitlines()[-1]);print('$1', 'value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e ms',round(d['e2e']['ms_per_step'],2),d['clocks']['sm_mhz'])
except Exception as e: print('$1 FAILED', e)"
}
run default FOO=1
run ctas8 NCCL_MAX_CTAS=8
run ctas4 NCCL_MAX_CTAS=4
run ctas16 NCCL_MAX_CTAS=16
run default2 FOO=1
