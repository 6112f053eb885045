#!/bin/bash
mkdir -p gpurun_out
P="python tools/probe_token.py 1 300"
for cfg in "13B 3 32" "13B 3 16" "13B 2 16" "7B 3 16"; do
  set -- $cfg
  echo "=== probe shape $1 type $2 slots $3"
  FASTLLAMA_B200_TK_SLOTS=$3 FASTLLAMA_B200_PROBE_SHAPE=$1 FASTLLAMA_B200_PROBE_TYPE=$2 timeout 100 $P > gpurun_out/tk12.txt 2>&1; echo "rc=$?"; grep -E "per launch|ERROR|rror" gpurun_out/tk12.txt | head -2
done
echo "=== 7B q4_0 probe, 8 layers"
timeout 100 python tools/probe_token.py 8 64 > gpurun_out/tk_prof12.txt 2>&1; echo "rc=$?"; grep -E "per launch|^ *(qkv|attn|wo|w13|w2|head):" gpurun_out/tk_prof12.txt | head -12
echo "=== token kernel vs CPU model, all shapes"
timeout 600 python -m pytest tests/test_gpu_fused.py -q -x -k "reference_bits" > gpurun_out/tok12.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/tok12.txt; grep -E "^E " gpurun_out/tok12.txt | head -8
