#!/bin/bash
mkdir -p gpurun_out
echo "=== quick regression of the committed binaries"
timeout 200 python -m pytest tests/test_golden_llama.py tests/test_gpu_e2e.py tests/test_gpu_graph.py -q -m gpu > gpurun_out/tests_last.txt 2>&1; echo "rc=$?"; tail -2 gpurun_out/tests_last.txt
timeout 120 python -m pytest tests/test_gpu_fused.py -q -k "reference_bits and (256 or 1408)" > gpurun_out/tests_last2.txt 2>&1; echo "rc=$?"; tail -2 gpurun_out/tests_last2.txt
echo "=== exact ingest (reference-order kernel for every eval)"
FASTLLAMA_B200_INGEST=exact timeout 200 python bench.py --mode ingest --steps 2 --no-extras > gpurun_out/bench_ingest_exact.json 2> gpurun_out/bench_ingest_exact.err; echo "rc=$?"; python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench_ingest_exact.json').read().strip().splitlines()[-1]); print("exact ingest value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"])
except Exception as e: print("no line", e); print(open('gpurun_out/bench_ingest_exact.err').read()[-1200:])
PY
