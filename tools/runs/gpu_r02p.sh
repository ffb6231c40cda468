#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02p}
mkdir -p $OUT
timeout 600 python tools/probe_misc.py posconv > $OUT/${TAG}_probe_posconv.log 2>&1; tail -15 $OUT/${TAG}_probe_posconv.log
timeout 1800 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log; tail -8 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench.json
timeout 600 python tools/profile_step.py > $OUT/${TAG}_profile_step.log 2>&1; head -30 $OUT/${TAG}_profile_step.log
