#!/bin/bash
mkdir -p gpurun_out
echo "=== umma"; timeout 600 python tools/probe_umma.py all > gpurun_out/umma_probe2.txt 2>&1; echo "rc=$?"; tail -28 gpurun_out/umma_probe2.txt
for d in 0 1; do
  echo "=== DIAG $d" >> gpurun_out/tk_prof2.txt
  FASTLLAMA_B200_TK_DIAG=$d timeout 120 python tools/probe_token.py 8 64 >> gpurun_out/tk_prof2.txt 2>&1
done
grep -E "DIAG|per launch|yfetch" gpurun_out/tk_prof2.txt
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "rc=$?"; tail -5 gpurun_out/bench2.err; cat gpurun_out/bench2.json
