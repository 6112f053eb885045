#!/bin/bash
mkdir -p gpurun_out
echo "=== f32 matmul / graph / golden / e2e tests"
timeout 300 python -m pytest tests/test_gpu_graph.py tests/test_golden_llama.py tests/test_gpu_e2e.py tests/test_lora.py -q -m gpu > gpurun_out/tests14.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/tests14.txt; grep -E "^E " gpurun_out/tests14.txt | head -8
echo "=== ingest bench"; timeout 300 python bench.py --mode ingest --steps 3 --no-extras > gpurun_out/bench_ingest14.json 2> gpurun_out/bench_ingest14.err; echo "rc=$?"; python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench_ingest14.json').read().strip().splitlines()[-1]); print("ingest value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"])
except Exception as e: print("no line", e); print(open('gpurun_out/bench_ingest14.err').read()[-1500:])
PY
echo "=== launch list, ingest"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_ingest.csv python bench.py --mode ingest --steps 1 --no-extras > gpurun_out/ncu_ingest.log 2>&1; echo "rc=$?"
