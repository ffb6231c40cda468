#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02w}
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 16 \
   -o $OUT/${TAG}_rowwise -f python tools/ncu_targets.py rowwise > $OUT/${TAG}_ncu_rowwise.log 2>&1
tail -2 $OUT/${TAG}_ncu_rowwise.log
