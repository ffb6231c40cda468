#!/bin/bash
# SpeechMixGAN: op-level + model-level parity on the GPU, with a diagnostic print-out of every number the test bounds
OUT=gpurun_out
TAG=${1:-r02aj}
mkdir -p $OUT
timeout 300 python tools/runs/diag_gan.py > $OUT/${TAG}_diag_gan.log 2>&1; grep -v Warning $OUT/${TAG}_diag_gan.log | tail -45
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -k "gan or gram" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log; grep -n "^E " $OUT/${TAG}_pytest.log | head -20
