"""CPU: every node of a LLaMA eval graph, reference library vs our host stack on the CPU stand-in of the device layer (tests/mock), bit
for bit.  Wider than the golden toy models (n_embd 1024, 8 heads of 128) and with a 45-token prompt, so that the attention products have
inner lengths with leftovers (45 = 32 + 8 + 4 + 1: every form of the compiled leftover loop of ggml_vec_dot_f32) and the quantised
matmuls run with 45 columns; then a decode step through the token program.  This is the check that located the last differences
(rope's fma contraction, the leftover forms) before the GPU kernels were written to the same order."""
import os
import subprocess
import sys

import pytest

from oracle.pyoracle import REF_GGML_SO
from tests.mockbuild import ensure_mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RUN = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
os.environ["FASTLLAMA_B200_SYNC_ALL"] = "1"            # copy every node's output back, not only the results
from oracle.pyoracle import REF_GGML_SO, Oracle
from tests import ggml_api as G
from tests.llama_graph import HParams, MiniLlama, make_weights
t = int(sys.argv[3])
orc = Oracle()
hp = HParams(n_vocab=96, n_embd=1024, n_head=8, n_layer=2, n_mult=256, n_ctx=64)
w = make_weights(hp, t, lambda x, tt: orc.quantize_q4(x, tt), seed=3)
models = [MiniLlama(G.Ggml(p), hp, w, compute_mb=256) for p in (REF_GGML_SO, sys.argv[2])]
for tokens, n_past, all_nodes in [(list(range(1, 46)), 0, True), ([7], 45, False), ([9], 46, False)]:
    res = []
    for m in models:
        c, gf, named = m.eval(tokens, n_past)
        m.compute(c, gf)
        vals = []
        for i in range(gf.n_nodes):
            tt = gf.nodes[i].contents
            inside = c.base <= tt.data < c.base + c.buf.nbytes
            vals.append((G.OP_NAMES[tt.op], c.numpy(gf.nodes[i]).copy() if inside and G.is_contiguous(tt) and tt.type == G.F32 else None))
        res.append(vals)
    pairs = list(zip(*res)) if all_nodes else [tuple(r[-1] for r in res)]       # a decode step only materialises its results
    for i, ((op, a), (_, b)) in enumerate(pairs):
        if a is None or b is None:
            continue
        nd = int((a.view(np.uint32) != b.view(np.uint32)).sum())
        assert nd == 0, (n_past, i, op, nd, a.size, float(np.abs(a - b).max()))
print("ok")
"""


@pytest.mark.skipif(not os.path.exists(REF_GGML_SO), reason="oracle/_ref not built")
@pytest.mark.parametrize("t", [2, 3])
def test_every_node_matches_the_reference_library(t):
    lib = os.path.join(ensure_mock(), "libggml_b200.so")
    if not os.path.exists(lib):
        pytest.skip("tests/mock not built")
    p = subprocess.run([sys.executable, "-c", RUN, ROOT, lib, str(t)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ok" in p.stdout, p.stderr[-3000:]
