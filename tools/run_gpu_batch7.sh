#!/bin/bash
mkdir -p gpurun_out
echo "=== token probe"
FASTLLAMA_B200_TK_DIAG=0 timeout 90 python tools/probe_token.py 8 64 > gpurun_out/tk_prof7.txt 2>&1; echo "rc=$?"
grep -E "per launch|^ *(qkv|attn|wo|w13|w2|head):" gpurun_out/tk_prof7.txt
if ! grep -q "per launch" gpurun_out/tk_prof7.txt; then tail -5 gpurun_out/tk_prof7.txt; echo "token kernel broken"; exit 1; fi
echo "=== depth probe"; timeout 400 python tools/probe_depth.py q4_0 7B 2>&1 | grep -E "^depth|Error|error" | tee gpurun_out/depth7.txt
echo "=== bench (with extras)"; timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "rc=$?"; grep "\[bench\]" gpurun_out/bench7.err | tail -6; cat gpurun_out/bench7.json
echo "=== full gpu tests"; timeout 1000 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_parity_full.py > gpurun_out/gputests7.txt 2>&1; tail -5 gpurun_out/gputests7.txt; grep -E "^E " gpurun_out/gputests7.txt | head -6
