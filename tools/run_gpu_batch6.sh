#!/bin/bash
mkdir -p gpurun_out
echo "=== token probe"
FASTLLAMA_B200_TK_DIAG=0 timeout 90 python tools/probe_token.py 8 64 > gpurun_out/tk_prof6.txt 2>&1; echo "rc=$?"
grep -E "per launch|^ *(qkv|attn|wo|w13|w2|head):" gpurun_out/tk_prof6.txt
if ! grep -q "per launch" gpurun_out/tk_prof6.txt; then tail -5 gpurun_out/tk_prof6.txt; echo "token kernel broken"; exit 1; fi
echo "=== token tests"; timeout 400 python -m pytest tests/test_gpu_fused.py -x -q -m gpu -k "token" > gpurun_out/tk_tests6.txt 2>&1; tail -3 gpurun_out/tk_tests6.txt; grep -E "^E " gpurun_out/tk_tests6.txt | head -4
echo "=== new gpu tests"; timeout 500 python -m pytest tests/test_lora.py tests/test_reload.py tests/test_gpu_rowfns.py -x -q -m gpu > gpurun_out/new_tests6.txt 2>&1; tail -3 gpurun_out/new_tests6.txt; grep -E "^E " gpurun_out/new_tests6.txt | head -6
echo "=== bench"; timeout 500 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "rc=$?"; grep "\[bench\]" gpurun_out/bench6.err | tail -4; cat gpurun_out/bench6.json
