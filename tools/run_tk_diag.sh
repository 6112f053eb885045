#!/bin/bash
# token-kernel timing decomposition (FASTLLAMA_B200_TK_DIAG bits: 1 no copies, 2 no dots, 4 no grid barriers, 8 no prologue, 16 no attention)
mkdir -p gpurun_out
out=gpurun_out/tk_diag.txt
: > $out
for d in 0 1 2 4 5 6 8 16 28 29 30; do
  echo "=== DIAG $d" >> $out
  FASTLLAMA_B200_TK_DIAG=$d timeout 120 python tools/probe_token.py 8 64 >> $out 2>&1
done
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -m gpu -k "token" > gpurun_out/tk_tests.txt 2>&1
tail -3 gpurun_out/tk_tests.txt
grep -E "DIAG|per launch|^ *(qkv|attn|wo|w13|w2|head):" $out
