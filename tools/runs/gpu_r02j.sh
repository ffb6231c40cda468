#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02j}
mkdir -p $OUT
timeout 600 python bench.py --steps 8 --warmup 3 > $OUT/${TAG}_bench_cfg2.json 2> $OUT/${TAG}_bench_cfg2.err; tail -c 1500 $OUT/${TAG}_bench_cfg2.json; tail -3 $OUT/${TAG}_bench_cfg2.err
for c in cfg3 cfg4 cfg5; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$c.json 2> $OUT/${TAG}_bench_$c.err; tail -c 1200 $OUT/${TAG}_bench_$c.json; tail -3 $OUT/${TAG}_bench_$c.err
done
timeout 300 python tools/profile_step.py > $OUT/${TAG}_profile_step.log 2>&1; head -30 $OUT/${TAG}_profile_step.log
