#!/bin/bash
# 2 GPUs: tensor-parallel tests + bench (row-split matrices, activation vectors gathered as dataflow words inside the token kernel)
mkdir -p gpurun_out
nvidia-smi -L
echo "=== tp tests"; [ -n "$SKIP_TP_TESTS" ] || timeout 400 python -m pytest tests/test_gpu_tp.py -x -q -m gpu > gpurun_out/tp_tests.txt 2>&1; tail -4 gpurun_out/tp_tests.txt; grep -E "^E " gpurun_out/tp_tests.txt | head -8
python bench.py --_gen 7B q4_0 > /dev/null 2>&1
echo "=== bench N=1 (for the tokens file + comparison on this box)"; timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/bench_tp_n1.json 2>/dev/null; python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_tp_n1.json').read().strip().splitlines()[-1]); print("N=1 value", j["value"], "e2e", j["e2e"]["value"])
PY
echo "=== bench N=2"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_tp2.json 2> gpurun_out/bench_tp2.err; echo "rc=$?"; python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench_tp2.json').read().strip().splitlines()[-1]); print("N=2 value", j["value"], "e2e", j["e2e"]["value"], "parity", json.dumps(j.get("parity"))[:600])
except Exception as e:
    print("no line", e); print(open('gpurun_out/bench_tp2.err').read()[-2000:])
PY
