#!/bin/bash
mkdir -p gpurun_out
P="python tools/probe_token.py 1 300"
echo "=== 13B q4_1 probe (bounded waits report what is stuck)"
FASTLLAMA_B200_PROBE_SHAPE=13B FASTLLAMA_B200_PROBE_TYPE=3 timeout 100 $P > gpurun_out/tk11_a.txt 2>&1; echo "rc=$?"; grep -E "per launch|ERROR|rror" gpurun_out/tk11_a.txt | head -3
echo "=== 13B q4_1, 8 slots"
FASTLLAMA_B200_TK_SLOTS=8 FASTLLAMA_B200_PROBE_SHAPE=13B FASTLLAMA_B200_PROBE_TYPE=3 timeout 100 $P > gpurun_out/tk11_b.txt 2>&1; echo "rc=$?"; grep -E "per launch|ERROR|rror" gpurun_out/tk11_b.txt | head -3
echo "=== 13B q4_0, 12 slots"
FASTLLAMA_B200_TK_SLOTS=12 FASTLLAMA_B200_PROBE_SHAPE=13B FASTLLAMA_B200_PROBE_TYPE=2 timeout 100 $P > gpurun_out/tk11_c.txt 2>&1; echo "rc=$?"; grep -E "per launch|ERROR|rror" gpurun_out/tk11_c.txt | head -3
echo "=== 7B q4_1, 12 slots"
FASTLLAMA_B200_TK_SLOTS=12 FASTLLAMA_B200_PROBE_TYPE=3 timeout 100 $P > gpurun_out/tk11_d.txt 2>&1; echo "rc=$?"; grep -E "per launch|ERROR|rror" gpurun_out/tk11_d.txt | head -3
echo "=== 7B q4_0 probe, 8 layers"
timeout 100 python tools/probe_token.py 8 64 > gpurun_out/tk_prof11.txt 2>&1; echo "rc=$?"; grep -E "per launch|^ *(qkv|attn|wo|w13|w2|head):" gpurun_out/tk_prof11.txt | head -8
