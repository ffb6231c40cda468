#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02n}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1800 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log; tail -8 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 8 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 1200 $OUT/${TAG}_bench.json
LIB_BATCH=32 timeout 600 python tools/bench_hf_gpu.py step > $OUT/${TAG}_library_step.log 2>&1; tail -2 $OUT/${TAG}_library_step.log | cut -c1-500
SMX_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
    --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
wc -l $OUT/${TAG}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 12 \
    -o $OUT/${TAG}_top -f python tools/ncu_targets.py attn gemm > $OUT/${TAG}_ncu_top.log 2>&1
tail -2 $OUT/${TAG}_ncu_top.log
