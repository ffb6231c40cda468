#!/bin/bash
# Fused dropout + residual + LayerNorm tail (forward and backward): parity, then the step with the recipe's dropout with
# and without the fusion on the same box; plus the deterministic model tests that walk through the edited autograd code
OUT=gpurun_out
TAG=${1:-r02ar}
mkdir -p $OUT
timeout 300 python -m pytest tests/test_dropout_gpu.py tests/test_model_gpu.py -q -m gpu -k "dropout or fused or eed_matches_oracle or mbart_pre_ln or adapter_matches_oracle or t5_v11" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log; grep -n "^E " $OUT/${TAG}_pytest.log | head -12
timeout 200 python bench.py --steps 8 --warmup 3 --dropout 0.1 --no-cpu-baseline > $OUT/${TAG}_bench_dropout.json 2> $OUT/${TAG}_bench_dropout.err; cut -c1-200 $OUT/${TAG}_bench_dropout.json; grep -i "error\|Traceback\|capture failed" $OUT/${TAG}_bench_dropout.err | head -3
SMX_DROPOUT_TAIL=0 timeout 200 python bench.py --steps 8 --warmup 3 --dropout 0.1 --no-cpu-baseline > $OUT/${TAG}_bench_dropout_unfused.json 2> /dev/null; cut -c1-200 $OUT/${TAG}_bench_dropout_unfused.json
