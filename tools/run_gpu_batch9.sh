#!/bin/bash
mkdir -p gpurun_out
echo "=== token probe (warp groups spread over the SM sub-partitions, unroll 8)"
FASTLLAMA_B200_TK_DIAG=0 timeout 90 python tools/probe_token.py 8 64 > gpurun_out/tk_prof9.txt 2>&1; echo "rc=$?"
grep -E "per launch|^ *(qkv|attn|wo|w13|w2|head):" gpurun_out/tk_prof9.txt
echo "=== shallow ring (12 slots, Sg = 3) on the small shapes"
FASTLLAMA_B200_TK_SLOTS=12 timeout 300 python -m pytest tests/test_gpu_fused.py -q -x -k "reference_bits and (256 or 1408)" > gpurun_out/tok9a.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/tok9a.txt; grep -E "^E " gpurun_out/tok9a.txt | head -5
echo "=== 13B q4_1 under compute-sanitizer"
timeout 900 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_fused.py -q -x -k "reference_bits and 5120 and 3]" > gpurun_out/tok9san.txt 2>&1; echo "rc=$?"; grep -E "Invalid|Error|error|at |by thread|Address|passed|failed" gpurun_out/tok9san.txt | head -30
echo "=== token kernel vs CPU model, all shapes"
timeout 600 python -m pytest tests/test_gpu_fused.py -q -k "reference_bits" > gpurun_out/tok9b.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/tok9b.txt; grep -E "^E " gpurun_out/tok9b.txt | head -8
echo "=== bench"; timeout 500 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "rc=$?"; grep "\[bench\]" gpurun_out/bench9.err | tail -3; python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench9.json').read().strip().splitlines()[-1])
print("value", j["value"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], "parity", j["parity"]["greedy_ids_equal"], j["parity"]["logits_maxabs_over_range"])
PY
