#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02r}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -k "adafactor or graph" > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 300 python tools/bench_adafactor.py > $OUT/${TAG}_adafactor.log 2>&1; tail -3 $OUT/${TAG}_adafactor.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --optimizer adafactor > $OUT/${TAG}_bench_adafactor.json 2> $OUT/${TAG}_bench_adafactor.err; cut -c1-300 $OUT/${TAG}_bench_adafactor.json; tail -3 $OUT/${TAG}_bench_adafactor.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench.json
