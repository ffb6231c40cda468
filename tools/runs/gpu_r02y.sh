#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02y}
mkdir -p $OUT
timeout 600 python tools/probe_misc.py conv0 > $OUT/${TAG}_probe_conv0.log 2>&1; tail -1 $OUT/${TAG}_probe_conv0.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:conv0 --csv --log-file $OUT/${TAG}_ncu_conv0.csv python tools/ncu_targets.py rowwise > /dev/null 2>&1; grep -v "^==" $OUT/${TAG}_ncu_conv0.csv | cut -d, -f5,13- | head -30
timeout 1800 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python -c "
import json;d=json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1]);print('graph ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],d['clocks'],d['roofline']['frac'])"
