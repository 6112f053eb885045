"""CPU: boundary B1.  libggml_b200 must be indistinguishable from the reference's ggml for everything
the host sees before ggml_graph_compute: struct layouts, arena accounting, builder results (shapes,
strides, data aliasing) and graph order.  Checked against the reference library itself
(oracle/_ref/libggml_ref.so) where it was built, and against layout constants otherwise.
No GPU compute is called here."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from fastllama_b200.build import lib_path
from oracle.pyoracle import REF_GGML_SO, REF_PYFASTLLAMA_SO, Oracle, have_ref
from tests import ggml_api as G
from tests.llama_graph import HParams, MiniLlama, graph_signature, make_weights

OURS = lib_path("libggml_b200.so")
pytestmark = pytest.mark.skipif(not os.path.exists(OURS), reason="native libs not built (run __graft_entry__.build())")


def _weights(hp, t=G.Q4_0):
    o = Oracle()
    return make_weights(hp, t, lambda w, tt: o.quantize_q4(w, tt))


def test_fp16_conversions_match_ieee():
    g = G.Ggml(OURS)
    bits = np.arange(1 << 16, dtype=np.uint16)
    ref32 = bits.view(np.float16).astype(np.float32)
    ours = np.array([g.fp16_to_fp32(int(b)) for b in bits[::7]], dtype=np.float32)
    want32 = ref32[::7]
    finite = ~np.isnan(want32)                      # NaN payloads do not survive a by-value float return
    assert np.array_equal(np.isnan(ours), np.isnan(want32))
    assert np.array_equal(ours[finite].view(np.uint32), want32[finite].view(np.uint32))
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.standard_normal(4000).astype(np.float32) * s for s in (1e-8, 1e-4, 1.0, 100.0, 7e4)])
    xs = np.concatenate([xs, np.array([0.0, -0.0, 65504.0, 65520.0, 65519.99, 6e-8, 2.98e-8, 2.99e-8, np.inf, -np.inf], dtype=np.float32)])
    with np.errstate(over="ignore"):
        want = xs.astype(np.float16).view(np.uint16)
    got = np.array([g.fp32_to_fp16(float(x)) for x in xs], dtype=np.uint16)
    assert np.array_equal(got, want)


def test_type_tables():
    g = G.Ggml(OURS)
    assert [g.type_size(t) for t in (0, 1, 2, 3, 6, 9)] == [4, 2, 20, 24, 40, 4]     # reference lib/ggml.c:3294-3307
    assert [g.blck_size(t) for t in (0, 1, 2, 3, 6, 9)] == [1, 1, 32, 32, 32, 1]
    assert [bool(g.is_quantized(t)) for t in (0, 1, 2, 3, 6, 9)] == [False, False, True, True, True, False]


def test_exported_symbols_cover_reference_imports():
    """Every ggml_* symbol the reference's upper layers import must be exported by libggml_b200, and
    the drop-in pyfastllama.so must export the reference's 17 llama_* entry points."""
    out = subprocess.run(["nm", "-D", "--defined-only", OURS], capture_output=True, text=True).stdout
    ours = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    needed = """ggml_add ggml_add_inplace ggml_blck_size ggml_build_forward ggml_build_forward_expand ggml_cpu_has_blas
    ggml_cpu_has_cublas ggml_cpy ggml_diag_mask_inf ggml_element_size ggml_fp16_to_fp32 ggml_free ggml_get_data
    ggml_get_rows ggml_graph_compute ggml_init ggml_is_quantized ggml_mul ggml_mul_mat ggml_nbytes ggml_nelements
    ggml_new_f32 ggml_new_tensor_1d ggml_new_tensor_2d ggml_permute ggml_quantize_chunk ggml_repeat ggml_reshape_2d
    ggml_reshape_3d ggml_rms_norm ggml_rope ggml_scale ggml_silu ggml_soft_max ggml_transpose ggml_type_name
    ggml_type_size ggml_used_mem ggml_view_1d ggml_view_2d ggml_view_3d ggml_internal_get_quantize_fn""".split()
    assert not [s for s in needed if s not in ours]
    dropin = lib_path("pyfastllama.so")
    if os.path.exists(dropin):
        out = subprocess.run(["nm", "-D", "--defined-only", dropin], capture_output=True, text=True).stdout
        syms = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
        want = """llama_create_default_context_args llama_create_context llama_load_model llama_set_stop_words llama_ingest
        llama_ingest_system_prompt llama_generate llama_perplexity llama_get_embeddings llama_get_logits llama_save_state
        llama_load_state llama_attach_lora llama_detach_lora llama_reset_model llama_free_context llama_handle_signal""".split()
        assert not [s for s in want if s not in syms]
        if os.path.exists(REF_PYFASTLLAMA_SO):
            out = subprocess.run(["nm", "-D", "--defined-only", REF_PYFASTLLAMA_SO], capture_output=True, text=True).stdout
            ref_llama = {ln.split()[-1] for ln in out.splitlines() if " T llama_" in ln}
            assert ref_llama <= syms


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n_tokens,n_past", [(1, 0), (1, 5), (4, 0), (3, 7)])
def test_eval_graph_identical_to_reference(n_tokens, n_past):
    """Same script, both libraries: identical node order, shapes, strides and arena offsets."""
    hp = HParams()
    w = _weights(hp)
    sigs = []
    for path in (REF_GGML_SO, OURS):
        g = G.Ggml(path)
        m = MiniLlama(g, hp, w, compute_mb=16)
        c, gf, named = m.eval(list(range(1, n_tokens + 1)), n_past)
        sig = graph_signature(c, gf, (m.wctx, m.kvctx))
        used = (g.used_mem(c.ctx), g.used_mem(m.wctx.ctx), g.used_mem(m.kvctx.ctx), gf.n_nodes, gf.n_leafs)
        sigs.append((sig, used))
    (sa, ua), (sb, ub) = sigs
    assert ua == ub
    assert sa[0] == sb[0]          # nodes
    assert sa[1] == sb[1]          # leafs


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_scratch_and_repeat_semantics_match_reference():
    for path in (REF_GGML_SO, OURS):
        g = G.Ggml(path)
        a = g.context(1 << 20)
        scratch = np.zeros(1 << 16, dtype=np.uint8)
        x = g.new_tensor_2d(a.ctx, G.F32, 8, 3)
        assert g.set_scratch(a.ctx, G.Scratch(0, scratch.nbytes, scratch.ctypes.data)) == 0
        y = g.new_tensor_2d(a.ctx, G.F32, 8, 3)                      # payload goes to the scratch buffer
        s = g.new_f32(a.ctx, 2.0)                                     # constants never do
        assert y.contents.data == scratch.ctypes.data
        assert a.base <= s.contents.data < a.base + (1 << 20)
        assert g.set_scratch(a.ctx, G.Scratch(0, 0, None)) == 8 * 3 * 4
        r = g.repeat(a.ctx, x, y)                                     # same shape -> returns its input
        assert C.addressof(r.contents) == C.addressof(x.contents)
        row = g.new_tensor_1d(a.ctx, G.F32, 8)
        r2 = g.repeat(a.ctx, row, y)
        assert G.OP_NAMES[r2.contents.op] == "REPEAT" and tuple(r2.contents.ne) == (8, 3, 1, 1)
        a.free()


def test_layout_constants_without_reference():
    assert C.sizeof(G.Tensor) == 176 and C.sizeof(G.CGraph) == 98360 and C.sizeof(G.InitParams) == 24
    g = G.Ggml(OURS)
    a = g.context(1 << 16)
    t = g.new_tensor_2d(a.ctx, G.Q4_0, 64, 3)
    # first object: 32-byte header, then the 176-byte tensor struct, payload right behind it
    assert a.offset(t) == 32 + 176 and tuple(t.contents.nb)[:2] == (20, 40)
    assert g.used_mem(a.ctx) == 32 + 176 + 128                       # 120 payload bytes rounded up to 16
    v = g.view_2d(a.ctx, t, 32, 3, 40, 20)
    assert v.contents.data == t.contents.data + 20 and tuple(v.contents.nb) == (20, 40, 120, 120)
    a.free()
