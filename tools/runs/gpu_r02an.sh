#!/bin/bash
# The recipe's dropout (0.1 at every site) inside the graphed step, and the SpeechMixGAN step at the cfg2 shapes (eager)
OUT=gpurun_out
TAG=${1:-r02an}
mkdir -p $OUT
timeout 300 python bench.py --steps 8 --warmup 3 --dropout 0.1 --no-cpu-baseline > $OUT/${TAG}_bench_dropout.json 2> $OUT/${TAG}_bench_dropout.err; cut -c1-330 $OUT/${TAG}_bench_dropout.json; grep -i "error\|Traceback\|capture failed" $OUT/${TAG}_bench_dropout.err | head -3
timeout 300 python bench.py --steps 8 --warmup 3 --config gan --no-graph --no-cpu-baseline > $OUT/${TAG}_bench_gan.json 2> $OUT/${TAG}_bench_gan.err; cut -c1-330 $OUT/${TAG}_bench_gan.json; grep -i "error\|Traceback" $OUT/${TAG}_bench_gan.err | head -3
timeout 300 python bench.py --steps 8 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_bench_eager.json 2> $OUT/${TAG}_bench_eager.err; cut -c1-330 $OUT/${TAG}_bench_eager.json
