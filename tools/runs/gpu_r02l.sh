#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02l}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_dropout_gpu.py -x -q -m gpu > $OUT/${TAG}_pytest_dropout.log 2>&1; tail -25 $OUT/${TAG}_pytest_dropout.log
