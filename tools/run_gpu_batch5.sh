#!/bin/bash
mkdir -p gpurun_out
echo "=== token probe"
FASTLLAMA_B200_TK_DIAG=0 timeout 90 python tools/probe_token.py 8 64 > gpurun_out/tk_prof5.txt 2>&1; echo "rc=$?"
grep -E "per launch|^ *(qkv|attn|wo|w13|w2|head):" gpurun_out/tk_prof5.txt
echo "=== token tests"; timeout 400 python -m pytest tests/test_gpu_fused.py -x -q -m gpu -k "token" > gpurun_out/tk_tests5.txt 2>&1; tail -3 gpurun_out/tk_tests5.txt; grep -E "^E " gpurun_out/tk_tests5.txt | head -8
echo "=== umma"; timeout 300 python tools/probe_umma.py all > gpurun_out/umma_probe5.txt 2>&1; echo "rc=$?"; grep -E "CHECK|FAIL|all quant|SYNC" gpurun_out/umma_probe5.txt | head -20
