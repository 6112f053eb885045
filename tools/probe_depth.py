"""GPU diagnostic: how does the logits difference between the reference (CPU, oracle/_ref) and this backend grow with DEPTH?
7B-width synthetic models truncated to 1, 2, 4, 8, 16 layers, same prompt, greedy, a few decode steps each.  A smooth growth is the
amplification of last-ulp differences (q8_0 rounding / fp16-table flips) through a random network; a jump at small depth would be a bug.

  python tools/probe_depth.py [q4_0|q4_1] [7B|13B]
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fastllama_b200.ggjt import write_synthetic_gpu  # noqa: E402

wt = sys.argv[1] if len(sys.argv) > 1 else "q4_0"
size = sys.argv[2] if len(sys.argv) > 2 else "7B"
be = bench.Backend(0)
N = 6
for depth in (1, 2, 4, 8, 16):
    path = os.path.join(bench.bench_dir(), f"depth_{size}_{wt}_{depth}.bin")
    write_synthetic_gpu(path, size=size, wtype={"q4_0": 2, "q4_1": 3}[wt], seed=0, std=0.02, n_layer=depth, fl=be.fl)
    with tempfile.TemporaryDirectory() as td:
        lp = os.path.join(td, "l.npy")
        r = bench.run_ref_worker({"path": path, "threads": 16, "prompt": bench.PROMPT, "n_parity": N, "logits_out": lp})
        ref_logits = np.load(lp)
    m = be.model(path, n_batch=1)
    assert m.ingest(bench.PROMPT)
    toks, logits = [], []
    for _ in range(N):
        got = []
        m.generate(lambda s: got.append(s), num_tokens=1, **bench.GREEDY)
        toks.append("".join(got))
        logits.append(m.get_logits_array())
    m.close()
    par = bench.compare_parity(r["parity_tokens"], ref_logits, toks, np.stack(logits))
    rel0 = float(np.abs(logits[0] - ref_logits[0]).max() / np.abs(ref_logits[0]).max())
    print(f"depth {depth:2d}: step-0 logits rel err {rel0:.3e}; over compared steps max {par.get('logits_maxabs_over_range', float('nan')):.3e} "
          f"median {par.get('logits_maxabs_over_range_median_step', float('nan')):.3e}; tokens equal {par['greedy_ids_equal']} (first divergence {par['first_divergence']}); "
          f"ref top1-top2 gap min {par.get('reference_top1_top2_gap_min', float('nan')):.3e}", flush=True)
    os.remove(path)
