#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02x}
mkdir -p $OUT
timeout 600 python tools/probe_misc.py conv0 > $OUT/${TAG}_probe_conv0.log 2>&1; grep -c '"nan": false' $OUT/${TAG}_probe_conv0.log; tail -2 $OUT/${TAG}_probe_conv0.log | cut -c1-200
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -q -m gpu -k "eed_matches or cfg1 or cfg2 or conv0 or spec or rowwise" > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -4 $OUT/${TAG}_pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 16 \
   -o $OUT/${TAG}_rowwise -f python tools/ncu_targets.py rowwise > $OUT/${TAG}_ncu_rowwise.log 2>&1
tail -1 $OUT/${TAG}_ncu_rowwise.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python -c "
import json;d=json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1]);print('graph ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],d['clocks'],d['roofline']['frac'])"
