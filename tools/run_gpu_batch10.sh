#!/bin/bash
mkdir -p gpurun_out
echo "=== 13B q4_1 token probe (1 layer), plain"
FASTLLAMA_B200_PROBE_SHAPE=13B FASTLLAMA_B200_PROBE_TYPE=3 timeout 120 python tools/probe_token.py 1 300 > gpurun_out/tk_13b_q41.txt 2>&1; rc=$?; echo "rc=$rc"; grep -E "per launch|Error|error" gpurun_out/tk_13b_q41.txt | head -3
if [ $rc -ne 0 ]; then
  echo "=== the same under compute-sanitizer"
  FASTLLAMA_B200_PROBE_SHAPE=13B FASTLLAMA_B200_PROBE_TYPE=3 timeout 240 compute-sanitizer --tool memcheck --print-limit 6 --kernel-name regex:k_decode_token python tools/probe_token.py 1 300 > gpurun_out/tk_13b_q41_san.txt 2>&1; echo "rc=$?"
  grep -E "Invalid|Error|at |by thread|Address|in block|misaligned|=====" gpurun_out/tk_13b_q41_san.txt | head -24
fi
echo "=== dataflow plan: golden toy model, graph evals, end to end, token kernel vs CPU model"
timeout 500 python -m pytest tests/test_golden_llama.py tests/test_gpu_graph.py tests/test_gpu_e2e.py tests/test_gpu_state.py tests/test_gpu_fused.py -q -m gpu -x --deselect "tests/test_gpu_fused.py::test_token_kernel_has_the_reference_bits[5120-40-13824-32000-512-300-1-3]" > gpurun_out/tests10.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/tests10.txt; grep -E "^E " gpurun_out/tests10.txt | head -12
echo "=== bench"; timeout 400 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "rc=$?"; grep "\[bench\]" gpurun_out/bench10.err | tail -3; python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench10.json').read().strip().splitlines()[-1])
    print("value", j["value"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], "parity", j["parity"]["greedy_ids_equal"], j["parity"].get("logits_bit_identical"), j["parity"]["logits_maxabs_over_range"])
except Exception as e: print("no bench line", e); print(open('gpurun_out/bench10.err').read()[-1500:])
PY
