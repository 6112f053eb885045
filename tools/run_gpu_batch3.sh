#!/bin/bash
mkdir -p gpurun_out
echo "=== token probe"
FASTLLAMA_B200_TK_DIAG=0 timeout 120 python tools/probe_token.py 8 64 > gpurun_out/tk_prof3.txt 2>&1
grep -E "per launch|^ *(qkv|attn|wo|w13|w2|head):" gpurun_out/tk_prof3.txt
echo "=== token tests"; timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -m gpu -k "token" 2>&1 | tail -3
echo "=== umma diag"
for d in 0 1 2 3; do echo "--- UMMA_DIAG $d"; FASTLLAMA_B200_UMMA_DIAG=$d timeout 300 python tools/probe_umma.py time 2>&1 | grep -E "impl [4567]:" | grep -E "output|w1|all quant"; done > gpurun_out/umma_diag.txt 2>&1
cat gpurun_out/umma_diag.txt
echo "=== bench"; timeout 700 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "rc=$?"; grep "\[bench\]" gpurun_out/bench3.err | tail; cat gpurun_out/bench3.json
echo "=== full gpu tests"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gputests3.txt 2>&1; tail -15 gpurun_out/gputests3.txt
