#!/bin/bash
# 2-GPU box: data-parallel correctness tests, then the driver's N=2 bench invocation (graph + eager roofline steps)
OUT=gpurun_out
TAG=${1:-r02s}
mkdir -p $OUT
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_parallel_gpu.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/${TAG}_pytest_2gpu.log
run() {  # config nproc steps
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $2 --config $1 --steps $3 --warmup 3 > $OUT/${TAG}_bench_$1_n$2.json 2> $OUT/${TAG}_bench_$1_n$2.err
  tail -c 1200 $OUT/${TAG}_bench_$1_n$2.json; grep -v Warning $OUT/${TAG}_bench_$1_n$2.err | tail -3 | cut -c1-300
}
run cfg2 2 8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $OUT/${TAG}_bench_ref_n2.json 2>/dev/null; cut -c1-300 $OUT/${TAG}_bench_ref_n2.json
