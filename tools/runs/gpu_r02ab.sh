#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02ab}
mkdir -p $OUT
timeout 600 python tools/probe_misc.py lmhead > $OUT/${TAG}_probe_lmhead.log 2>&1; tail -2 $OUT/${TAG}_probe_lmhead.log | cut -c1-200
timeout 1800 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python -c "
import json;d=json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1]);print('graph ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],d['clocks'],d['roofline']['frac'])"
timeout 600 python tools/profile_step.py > $OUT/${TAG}_profile_step.log 2>&1; grep -n "k8192\|step_ms\|conv0" $OUT/${TAG}_profile_step.log
