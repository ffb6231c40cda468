#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -c 6 \
    -o $OUT/r02c_attn -f python tools/ncu_targets.py attn > $OUT/r02c_ncu_attn.log 2>&1
tail -3 $OUT/r02c_ncu_attn.log
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -q -m gpu > $OUT/r02c_pytest_fullsize.log 2>&1; tail -15 $OUT/r02c_pytest_fullsize.log
