#!/bin/bash
# first GPU contact of the reference-order kernels: row functions, token kernel against the CPU model, golden toy model, then timing
mkdir -p gpurun_out
echo "=== row functions (general reference-order kernel)"
timeout 400 python -m pytest tests/test_gpu_rowfns.py -q -x > gpurun_out/rowfns8.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/rowfns8.txt; grep -E "^E " gpurun_out/rowfns8.txt | head -8
echo "=== token kernel vs CPU model, small shapes"
timeout 300 python -m pytest tests/test_gpu_fused.py -q -x -k "reference_bits and (256 or 1408)" > gpurun_out/tok8a.txt 2>&1; rc=$?; echo "rc=$rc"; tail -3 gpurun_out/tok8a.txt; grep -E "^E " gpurun_out/tok8a.txt | head -8
echo "=== token kernel vs CPU model, 7B / 13B shapes"
timeout 600 python -m pytest tests/test_gpu_fused.py -q -k "reference_bits and not (256 or 1408)" > gpurun_out/tok8b.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/tok8b.txt; grep -E "^E " gpurun_out/tok8b.txt | head -8
echo "=== golden toy model + graph ops"
timeout 400 python -m pytest tests/test_golden_llama.py tests/test_gpu_graph.py tests/test_lora.py -q -m gpu > gpurun_out/graph8.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/graph8.txt; grep -E "^E " gpurun_out/graph8.txt | head -12
echo "=== token probe"
FASTLLAMA_B200_TK_DIAG=0 timeout 90 python tools/probe_token.py 8 64 > gpurun_out/tk_prof8.txt 2>&1; echo "rc=$?"
grep -E "per launch|^ *(qkv|attn|wo|w13|w2|head):" gpurun_out/tk_prof8.txt
if ! grep -q "per launch" gpurun_out/tk_prof8.txt; then tail -5 gpurun_out/tk_prof8.txt; echo "token kernel broken"; exit 1; fi
echo "=== bench"; timeout 500 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "rc=$?"; grep "\[bench\]" gpurun_out/bench8.err | tail -4; cat gpurun_out/bench8.json
