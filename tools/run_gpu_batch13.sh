#!/bin/bash
mkdir -p gpurun_out
P="python tools/probe_token.py 1 300"
for cfg in "13B 3 64" "13B 3 8" "13B 2 12" "7B 3 12"; do
  set -- $cfg
  echo "=== probe shape $1 type $2 slots $3"
  FASTLLAMA_B200_TK_SLOTS=$3 FASTLLAMA_B200_PROBE_SHAPE=$1 FASTLLAMA_B200_PROBE_TYPE=$2 timeout 100 $P > gpurun_out/tk13.txt 2>&1; echo "rc=$?"; grep -E "per launch|ERROR|rror" gpurun_out/tk13.txt | head -2
done
echo "=== 7B q4_0 probe, 8 layers"
timeout 100 python tools/probe_token.py 8 64 > gpurun_out/tk_prof13.txt 2>&1; echo "rc=$?"; grep -E "per launch|^ *(qkv|attn|wo|w13|w2|head):" gpurun_out/tk_prof13.txt | head -7
echo "=== token kernel vs CPU model, all shapes"
timeout 600 python -m pytest tests/test_gpu_fused.py -q -x -k "reference_bits" > gpurun_out/tok13.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/tok13.txt; grep -E "^E " gpurun_out/tok13.txt | head -8
echo "=== bench"; timeout 400 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "rc=$?"; python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench13.json').read().strip().splitlines()[-1])
    print("value", j["value"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], "parity", j["parity"]["greedy_ids_equal"], j["parity"].get("logits_bit_identical"), j["parity"]["logits_maxabs_over_range"])
except Exception as e: print("no bench line", e); print(open('gpurun_out/bench13.err').read()[-1500:])
PY
