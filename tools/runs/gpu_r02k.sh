#!/bin/bash
# 8-GPU box: 2-GPU correctness tests of the data-parallel modes, then scaling runs of the BASELINE configs
OUT=gpurun_out
TAG=${1:-r02k}
mkdir -p $OUT
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_parallel_gpu.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/${TAG}_pytest_2gpu.log
run() {  # config nproc steps
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $2 --config $1 --steps $3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$1_n$2.json 2> $OUT/${TAG}_bench_$1_n$2.err
  tail -c 900 $OUT/${TAG}_bench_$1_n$2.json; tail -2 $OUT/${TAG}_bench_$1_n$2.err | cut -c1-300
}
run cfg2 8 8
run cfg3 2 5
run cfg3 8 5
run cfg5 8 5
