#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02t}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1800 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log; tail -8 $OUT/${TAG}_pytest_gpu.log
timeout 300 python tools/micro/hbm_rw.py > $OUT/${TAG}_hbm_rw.txt 2>&1; cat $OUT/${TAG}_hbm_rw.txt
timeout 600 python bench.py --steps 8 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench.json
timeout 900 python -m pytest tests -q -m gpu -x > $OUT/${TAG}_pytest_gpu_second.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu_second.log; tail -4 $OUT/${TAG}_pytest_gpu_second.log
