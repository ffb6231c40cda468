#!/bin/bash
# 8-GPU box, final code: 2-GPU correctness tests + the driver's N = 8 invocation of both arms
OUT=gpurun_out
TAG=${1:-r02ag}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_parallel_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee $OUT/${TAG}_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --impl reference --gpus 8 --steps 1 --warmup 1 > $OUT/${TAG}_bench_ref_n8.json 2> $OUT/${TAG}_bench_ref_n8.err; cut -c1-200 $OUT/${TAG}_bench_ref_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 8 --warmup 3 > $OUT/${TAG}_bench_cfg2_n8.json 2> $OUT/${TAG}_bench_cfg2_n8.err
python -c "
import json
d=json.loads(open('$OUT/${TAG}_bench_cfg2_n8.json').read().strip().splitlines()[-1]);print('cfg2 N=8 value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']),d['clocks'])"
grep -v Warning $OUT/${TAG}_bench_cfg2_n8.err | grep -i "error\|Traceback" | head -3
