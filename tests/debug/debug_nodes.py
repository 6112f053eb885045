"""Debug aid: run the mini LLaMA graph on the reference (CPU) and on libggml_b200 (GPU) and print
the first nodes whose outputs diverge."""
import os
import sys

import numpy as np

os.environ["FASTLLAMA_B200_SYNC_ALL"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fastllama_b200.build import lib_path  # noqa: E402
from oracle.pyoracle import REF_GGML_SO, Oracle  # noqa: E402
from tests import ggml_api as G  # noqa: E402
from tests.llama_graph import HParams, MiniLlama, make_weights  # noqa: E402

t = int(sys.argv[1]) if len(sys.argv) > 1 else G.Q4_1
orc = Oracle()
hp = HParams(n_vocab=96, n_embd=128, n_head=4, n_layer=3, n_mult=32, n_ctx=32)
w = make_weights(hp, t, lambda x, tt: orc.quantize_q4(x, tt), seed=3)
libs = [G.Ggml(REF_GGML_SO), G.Ggml(lib_path("libggml_b200.so"))]
models = [MiniLlama(g, hp, w, compute_mb=32) for g in libs]
for tokens, n_past in [([5, 17, 3, 80, 41], 0), ([7], 5)]:
    res = []
    for m in models:
        c, gf, named = m.eval(tokens, n_past)
        m.compute(c, gf)
        vals = []
        for i in range(gf.n_nodes):
            tt = gf.nodes[i].contents
            inside = c.base <= tt.data < c.base + c.buf.nbytes
            vals.append((G.OP_NAMES[tt.op], tuple(tt.ne), c.numpy(gf.nodes[i]).copy() if inside and G.is_contiguous(tt) and tt.type == G.F32 else None))
        res.append(vals)
    print(f"--- tokens={tokens} n_past={n_past}")
    shown = 0
    for i, ((op, ne, a), (_, _, b)) in enumerate(zip(*res)):
        if a is None or b is None:
            continue
        d = np.abs(a - b).max()
        rel = d / max(np.abs(a).max(), 1e-30)
        if rel > 1e-6 or not np.isfinite(b).all():
            print(f"node {i:4d} {op:14s} ne={ne} max|ref|={np.abs(a).max():.4g} maxdiff={d:.3g} rel={rel:.3g}")
            shown += 1
            if shown > 25:
                break
