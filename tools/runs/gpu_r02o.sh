#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02o}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1800 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log; tail -8 $OUT/${TAG}_pytest_gpu.log
SMX_PDL=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_nopdl.json 2> $OUT/${TAG}_bench_nopdl.err; cut -c1-400 $OUT/${TAG}_bench_nopdl.json
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-400 $OUT/${TAG}_bench.json
SMX_PDL=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_nopdl2.json 2> $OUT/${TAG}_bench_nopdl2.err; cut -c1-400 $OUT/${TAG}_bench_nopdl2.json
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench2.json 2> $OUT/${TAG}_bench2.err; cut -c1-400 $OUT/${TAG}_bench2.json
timeout 300 python tools/bench_adafactor.py > $OUT/${TAG}_adafactor.log 2>&1; tail -5 $OUT/${TAG}_adafactor.log
