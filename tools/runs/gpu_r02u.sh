#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02u}
mkdir -p $OUT
timeout 600 python tools/probe_misc.py conv0 > $OUT/${TAG}_probe_conv0.log 2>&1; tail -8 $OUT/${TAG}_probe_conv0.log | cut -c1-200
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -q -m gpu -k "graphed or beam or conv0 or eed_matches or cfg1" > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -4 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench.json; python -c "
import json;d=json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1]);print('ms',d['ms_per_step'],'e2e',d['e2e'])"
timeout 600 python tools/profile_step.py > $OUT/${TAG}_profile_step.log 2>&1; grep -n "conv0\|step_ms" $OUT/${TAG}_profile_step.log
