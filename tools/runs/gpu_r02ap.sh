#!/bin/bash
# Packed-pair dropout paths of the attention backward kernels: parity (tests/test_dropout_gpu.py), then the step with the
# recipe's dropout (0.1 at every site) -- bench line and per-kernel profile
OUT=gpurun_out
TAG=${1:-r02ap}
mkdir -p $OUT
timeout 400 python -m pytest tests/test_dropout_gpu.py -q -m gpu > $OUT/${TAG}_pytest_dropout.log 2>&1; tail -3 $OUT/${TAG}_pytest_dropout.log; grep -n "^E " $OUT/${TAG}_pytest_dropout.log | head -12
timeout 300 python bench.py --steps 8 --warmup 3 --dropout 0.1 --no-cpu-baseline > $OUT/${TAG}_bench_dropout.json 2> $OUT/${TAG}_bench_dropout.err; cut -c1-330 $OUT/${TAG}_bench_dropout.json; grep -i "error\|Traceback\|capture failed" $OUT/${TAG}_bench_dropout.err | head -3
timeout 200 python tools/profile_step.py 32 0.1 > $OUT/${TAG}_profile_dropout.log 2>&1; grep -v Warn $OUT/${TAG}_profile_dropout.log | sed -n 3,12p
