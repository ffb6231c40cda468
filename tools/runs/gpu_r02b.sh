#!/bin/bash
# round-2 call B: new attention forward (correctness + perf A/B + SDPA baseline), then the new parity tests
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r02b_smi.txt 2>&1
timeout 600 python tools/probe_attn.py fwd2_rescale fwd_small kv_len_mask full_size > $OUT/r02b_probe_attn.log 2>&1
grep -c '"rel"' $OUT/r02b_probe_attn.log; grep "case_done\|rc=" $OUT/r02b_probe_attn.log
for v in "SMX_ATTN_FWD_V1=1" "SMX_ATTN_POLY=0" "SMX_ATTN_POLY=4" "SMX_ATTN_POLY=3"; do
  env $v timeout 300 python tools/probe_attn.py --case perf 2>&1 | grep perf >> $OUT/r02b_attn_perf.log
done
cat $OUT/r02b_attn_perf.log
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -x -q -m gpu > $OUT/r02b_pytest_fullsize.log 2>&1; tail -15 $OUT/r02b_pytest_fullsize.log
