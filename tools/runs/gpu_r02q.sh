#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02q}
mkdir -p $OUT
for pdl in 0 1 0 1 0 1; do
  SMX_PDL=$pdl timeout 600 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -k "cfg3 or cfg5" > $OUT/${TAG}_cfg3_pdl${pdl}.log 2>&1
  python - <<PY
import json
for c in ("cfg3_adapter_hubert_large_bart_large","cfg5_eed_hubert_large_mbart50"):
    d=json.load(open("gpurun_out/parity_%s.json"%c))
    print("pdl=$pdl", c, "loss %.6f dloss %.6f flips %d logits_rel %.5f speech_rel %.5f"%(d["loss"],d["dloss"],d["id_flips"],d["logits_rel_max"],d["speech_rel_max"]))
PY
done | tee $OUT/${TAG}_pdl_determinism.txt
timeout 300 python tools/bench_adafactor.py > $OUT/${TAG}_adafactor.log 2>&1; tail -4 $OUT/${TAG}_adafactor.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -5 $OUT/${TAG}_pytest_gpu.log
