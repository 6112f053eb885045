#!/bin/bash
# round-2 closing run on one GPU: bench line (with extras), ncu evidence for profiles/, the whole -m gpu suite
mkdir -p gpurun_out
echo "=== bench (with extras)"; timeout 800 python bench.py --steps 32 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; grep "\[bench\]" gpurun_out/bench_final.err | tail -5
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
    print("value", j["value"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], "parity", j["parity"]["greedy_ids_equal"], j["parity"].get("logits_bit_identical"))
    for e in j.get("extra", []): print("  extra:", e["metric"], round(e["value"],1), "e2e", round(e["e2e"]["value"],1), "frac", round(e["roofline"]["frac"],4), e.get("parity", {}).get("greedy_ids_equal") if isinstance(e.get("parity"), dict) else "")
except Exception as ex: print("no bench line", ex); print(open('gpurun_out/bench_final.err').read()[-1500:])
PY
B="python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline"
echo "=== launch list, decode"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_decode.csv $B > gpurun_out/ncu_decode.log 2>&1; echo "rc=$?"
echo "=== full capture, token kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_decode_token -s 8 -c 1 -o gpurun_out/r02_token_exact -f $B > gpurun_out/ncu_token.log 2>&1; echo "rc=$?"
echo "=== full gpu tests"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/gputests_final.txt 2>&1; echo "rc=$?"; tail -6 gpurun_out/gputests_final.txt; grep -E "^E |^FAILED" gpurun_out/gputests_final.txt | head -12
echo "=== launch list, ingest"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/r02_launches_ingest.csv python bench.py --mode ingest --steps 1 > gpurun_out/ncu_ingest.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_*.csv
