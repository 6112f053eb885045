#!/bin/bash
# 2 GPUs: tensor-parallel tests + bench (LL reductions, then the round-1 flag-barrier form for comparison)
mkdir -p gpurun_out
nvidia-smi -L
echo "=== tp tests"; timeout 600 python -m pytest tests/test_gpu_tp.py -x -q -m gpu > gpurun_out/tp_tests.txt 2>&1; tail -4 gpurun_out/tp_tests.txt; grep -E "^E " gpurun_out/tp_tests.txt | head -8
python bench.py --_gen 7B q4_0 > /dev/null 2>&1
echo "=== bench N=1 (for the tokens file + comparison on this box)"; timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/bench_tp_n1.json 2>/dev/null; cat gpurun_out/bench_tp_n1.json | head -c 600; echo
echo "=== bench N=2 (LL)"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_tp2_ll.json 2> gpurun_out/bench_tp2_ll.err; echo "rc=$?"; cat gpurun_out/bench_tp2_ll.json; grep -iE "error|fail|timeout" gpurun_out/bench_tp2_ll.err | head -5
echo "=== bench N=2 (flag barrier)"; FASTLLAMA_B200_TP_NO_LL=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_tp2_flags.json 2> gpurun_out/bench_tp2_flags.err; echo "rc=$?"; cat gpurun_out/bench_tp2_flags.json
