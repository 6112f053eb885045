#!/bin/bash
mkdir -p gpurun_out
echo "=== token probe"
FASTLLAMA_B200_TK_DIAG=0 timeout 90 python tools/probe_token.py 8 64 > gpurun_out/tk_prof4.txt 2>&1; echo "rc=$?"
grep -E "per launch|^ *(qkv|attn|wo|w13|w2|head):" gpurun_out/tk_prof4.txt
if grep -q "per launch" gpurun_out/tk_prof4.txt; then
  echo "=== token tests"; timeout 400 python -m pytest tests/test_gpu_fused.py -x -q -m gpu -k "token" 2>&1 | tail -3
else
  tail -5 gpurun_out/tk_prof4.txt; echo "token kernel broken: skipping token tests and bench"
fi
echo "=== umma"; timeout 300 python tools/probe_umma.py all > gpurun_out/umma_probe4.txt 2>&1; echo "rc=$?"; grep -E "CHECK|FAIL|all quant|SYNC" gpurun_out/umma_probe4.txt | head -20
if grep -q "per launch" gpurun_out/tk_prof4.txt; then
  echo "=== bench"; timeout 500 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "rc=$?"; grep "\[bench\]" gpurun_out/bench4.err | tail; cat gpurun_out/bench4.json
fi
