#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02ad}
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -q -m gpu -k "t5 or env_selected or no_grad_forward or beam or conv0 or self_matches" > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -25 $OUT/${TAG}_pytest_gpu.log | cut -c1-300
