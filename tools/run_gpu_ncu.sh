#!/bin/bash
# ncu evidence for profiles/ (one GPU; numbers printed under ncu are never bench values)
mkdir -p gpurun_out
python bench.py --_gen 7B q4_0 > /dev/null 2>&1
B="python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline"
echo "=== launch list, decode"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_decode.csv $B > gpurun_out/ncu_decode.log 2>&1; echo "rc=$?"
echo "=== full capture, token kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_token -s 8 -c 1 -o gpurun_out/r02_token -f $B > gpurun_out/ncu_token.log 2>&1; echo "rc=$?"
echo "=== launch list, ingest"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/r02_launches_ingest.csv python bench.py --mode ingest --steps 1 > gpurun_out/ncu_ingest.log 2>&1; echo "rc=$?"
echo "=== full capture, tcgen05 GEMM (w1-sized and output-sized launches)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mul_mat_q_umma -s 228 -c 3 -o gpurun_out/r02_umma -f python bench.py --mode ingest --steps 1 > gpurun_out/ncu_umma.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_*.csv
echo "=== ingest bench line (no profiler)"; timeout 300 python bench.py --mode ingest --steps 2 2> /dev/null | tee gpurun_out/bench_ingest.json
