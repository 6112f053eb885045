"""Generate tests/golden/*.npz from the REFERENCE itself (oracle/_ref/libggml_ref.so).

Run in the build container (where /root/reference exists):  python -m oracle.gen_golden
The reference has no golden vectors of its own for this path, so these
fixtures are outputs of the reference's row kernels on seeded inputs, produced through the
hook the reference exports for tests (ggml_internal_get_quantize_fn, include/ggml.h:841-862).
They pin the C restatement (tests/test_oracle.py) and the CUDA path (tests/test_gpu_*.py) on
machines where only the committed files exist.
"""
import os

import numpy as np

from oracle.pyoracle import GGML_TYPE_Q4_0, GGML_TYPE_Q4_1, RefGgml, build_oracle

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def activation_rows(rng, n, k):
    """Seeded activations with the edge cases the q8_0 quantizer has: all-zero block, single
    non-zero, exact .5 ties after scaling (amax = 127 -> id = 1), denormal and huge scales,
    negative amax, constant blocks."""
    x = rng.standard_normal((n, k)).astype(np.float32)
    x[0] = 0.0
    x[1] = 0.0
    x[1, 5] = 1.0
    x[2, :32] = np.arange(32, dtype=np.float32) - 15.5
    x[2, 0] = 127.0
    x[3, :32] = np.linspace(-63.5, 63.5, 32, dtype=np.float32)
    x[3, 31] = -127.0
    x[4] *= np.float32(1e-30)
    x[5] *= np.float32(1e30)
    x[6] = np.round(x[6] * 4) / 4
    x[7, 32:64] = 3.0
    return x


LLAMA_TOY = dict(n_vocab=96, n_embd=256, n_head=4, n_layer=3, n_mult=256, n_ctx=32)
LLAMA_TOY_STEPS = [([5, 17, 3, 80, 41], 0), ([7], 5), ([60], 6), ([2], 7)]
LLAMA_TOY_SEED = 3


def llama_toy(ref):
    """Whole-graph fixture: logits and final embeddings of the reference LIBRARY (libggml_ref.so, CPU) for a tiny LLaMA
    built through the ggml C API exactly like Model::eval builds it (tests/llama_graph.py): a 5-token prompt, then three
    decode steps that read the KV cache.  Weights come from tests.llama_graph.make_weights (numpy seed + the reference's
    own file quantiser), so a test on a machine without /root/reference rebuilds the same model bit for bit."""
    from oracle.pyoracle import REF_GGML_SO, Oracle
    from tests import ggml_api as G
    from tests.llama_graph import HParams, MiniLlama, make_weights

    orc = Oracle()
    g = G.Ggml(REF_GGML_SO)
    for name, t in (("q4_0", G.Q4_0), ("q4_1", G.Q4_1)):
        hp = HParams(**LLAMA_TOY)
        w = make_weights(hp, t, lambda x, tt: orc.quantize_q4(x, tt), seed=LLAMA_TOY_SEED)
        m = MiniLlama(g, hp, w, compute_mb=32)
        out = {}
        for i, (tokens, n_past) in enumerate(LLAMA_TOY_STEPS):
            c, gf, named = m.eval(tokens, n_past)
            m.compute(c, gf)
            out[f"logits{i}"] = c.numpy(named["logits"]).copy()
            out[f"emb{i}"] = c.numpy(named["embeddings"]).copy()
        path = os.path.join(OUT, f"llama_toy_{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


LORA_SHAPE = dict(k=256, m=24, ranks=(8, 16, 40, 64))


def lora_inputs(rng, m, k):
    """Rows for the SIMD re-quantisers and the LoRA merge: random, ties after scaling (amax = 7 -> id = 1: x.5 values),
    all-zero and constant blocks, a block whose merged values change sign."""
    w = (rng.standard_normal((m, k)) * 0.05).astype(np.float32)
    w[0] = 0.0
    w[1, :32] = 0.125
    w[2, :32] = (np.arange(32, dtype=np.float32) - 15.5) * 0.4375        # amax = 6.78..: scaled values land near .5 ties
    w[3, :32] = np.linspace(-7.0, 7.0, 32, dtype=np.float32)
    w[3, 0:8] = [0.5, 1.5, 2.5, 3.5, -0.5, -1.5, -2.5, -3.5]
    w[4, :32] = np.linspace(0.0, 15.0, 32, dtype=np.float32)
    w[4, 1:5] = [0.5, 1.5, 2.5, 14.5]                                      # min 0, max 15: d = id = 1, exact ties
    return w


def lora_ops():
    """Outputs of the reference LIBRARY for the ops of attach_lora / detach_lora (lib/llama.cpp:697-944): the SIMD row
    quantisers (quantize_fns[].quantize_row_q), f32 x f32 mul_mat (B*A) and add_inplace of an f32 matrix into a quantised one."""
    from oracle.pyoracle import REF_GGML_SO
    from tests import ggml_api as G

    ref = RefGgml()
    g = G.Ggml(REF_GGML_SO)
    rng = np.random.default_rng(77)
    k, m = LORA_SHAPE["k"], LORA_SHAPE["m"]
    w = lora_inputs(rng, m, k)
    out = {"w": w}
    for name, t in (("q4_0", GGML_TYPE_Q4_0), ("q4_1", GGML_TYPE_Q4_1)):
        out[f"{name}_simd"] = ref.quantize_q4_simd(w, t)
        base = ref.quantize_q4_reference(w, t)
        out[f"{name}_base"] = base
    for r in LORA_SHAPE["ranks"]:
        a = (rng.standard_normal((k, r)) * 0.3).astype(np.float32)      # loraA: ne = [r, k]
        b = (rng.standard_normal((m, r)) * 0.3).astype(np.float32)      # loraB: ne = [r, m]
        out[f"A{r}"], out[f"B{r}"] = a, b
        for name, t in (("q4_0", G.Q4_0), ("q4_1", G.Q4_1)):
            ar = g.context(64 << 20)
            ta = g.new_tensor_2d(ar.ctx, G.F32, r, k); ar.set(ta, a)
            tb = g.new_tensor_2d(ar.ctx, G.F32, r, m); ar.set(tb, b)
            tw = g.new_tensor_2d(ar.ctx, t, k, m); ar.set(tw, out[f"{name}_base"])
            ba = g.mul_mat(ar.ctx, ta, tb)
            res = g.add_inplace(ar.ctx, tw, ba)
            gf = G.new_graph()
            g.build_forward_expand(gf, res)
            g.graph_compute(ar.ctx, gf)
            if name == "q4_0":
                out[f"BA{r}"] = ar.numpy(ba).reshape(m, k).copy()
            out[f"{name}_merged{r}"] = ar.numpy(tw).reshape(m, -1).copy()
            # detach: W - BA (scale by -1, add in place)
            neg = g.scale(ar.ctx, ba, g.new_f32(ar.ctx, -1.0))
            res2 = g.add_inplace(ar.ctx, tw, neg)
            gf2 = G.new_graph()
            g.build_forward_expand(gf2, res2)
            g.graph_compute(ar.ctx, gf2)
            out[f"{name}_detached{r}"] = ar.numpy(tw).reshape(m, -1).copy()
            ar.free()
    path = os.path.join(OUT, "lora_ops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


F32_DOT_KS = (1, 3, 4, 5, 7, 8, 9, 12, 13, 15, 16, 31, 32, 33, 36, 37, 39, 44, 45, 47, 63, 64, 77, 128, 200)


def f32_dot():
    """f32 x f32 mul_mat of the reference LIBRARY for inner lengths with every kind of leftover count (n % 32): pins how the compiled
    reference adds the leftovers of ggml_vec_dot_f32 (lib/ggml.c:2316-2319; gcc vectorises that loop, see oracle/q4_oracle.c)."""
    from oracle.pyoracle import REF_GGML_SO
    from tests import ggml_api as G

    g = G.Ggml(REF_GGML_SO)
    rng = np.random.default_rng(4242)
    out = {}
    m, n = 6, 5
    for k in F32_DOT_KS:
        a = rng.standard_normal((m, k)).astype(np.float32)
        b = rng.standard_normal((n, k)).astype(np.float32)
        ar = g.context(4 << 20)
        ta = g.new_tensor_2d(ar.ctx, G.F32, k, m); ar.set(ta, a)
        tb = g.new_tensor_2d(ar.ctx, G.F32, k, n); ar.set(tb, b)
        r = g.mul_mat(ar.ctx, ta, tb)
        gf = G.new_graph()
        g.build_forward_expand(gf, r)
        g.graph_compute(ar.ctx, gf)
        out[f"a{k}"], out[f"b{k}"], out[f"out{k}"] = a, b, ar.numpy(r).reshape(n, m).copy()
        ar.free()
    path = os.path.join(OUT, "f32_dot.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    build_oracle()
    ref = RefGgml()
    rng = np.random.default_rng(20230423)
    os.makedirs(OUT, exist_ok=True)

    # K = 64 is the smallest legal row (vec_dot asserts an even block count, lib/ggml.c:2372)
    for k, m, n in ((64, 8, 8), (256, 40, 9), (4096, 6, 8)):
        x = activation_rows(rng, n, k)
        w = (rng.standard_normal((m, k)) * 0.02).astype(np.float32)
        w[0] = 0.0                      # all-zero weight row: d = 0 blocks
        w[1, :32] = 0.02                # constant block: q4_1 d = 0, m = value
        out = {"x": x, "w": w, "q8": ref.quantize_q8_0(x)}
        for name, t in (("q4_0", GGML_TYPE_Q4_0), ("q4_1", GGML_TYPE_Q4_1)):
            wq = ref.quantize_q4_reference(w, t)
            out[f"{name}_w"] = wq
            out[f"{name}_deq"] = ref.dequantize_q4(wq, t, k)
            out[f"{name}_mul_mat"] = ref.mul_mat_q(wq, x, t)        # [N, M]
        path = os.path.join(OUT, f"rowfns_k{k}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")
    llama_toy(ref)
    lora_ops()
    f32_dot()


if __name__ == "__main__":
    main()
