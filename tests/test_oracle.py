"""CPU: pin the C restatement (oracle/q4_oracle.c) to the reference.

Two pins, as the reference itself ships no tests for this path:
  1. the committed fixtures tests/golden/rowfns_k*.npz -- outputs of the reference's own row
     kernels (oracle/gen_golden.py), available everywhere;
  2. the live reference library oracle/_ref/libggml_ref.so on fresh random inputs, wherever it
     was built.
Everything here is integer/byte work or a fixed fp32 operation order, so the bar is bit-exact.
"""
import numpy as np
import pytest

from oracle.pyoracle import (GGML_TYPE_Q4_0, GGML_TYPE_Q4_1, Q4_0_DT, Q4_1_DT, Q8_0_DT, np_quantize_q4_0,
                             np_quantize_q4_1)

TYPES = [("q4_0", GGML_TYPE_Q4_0), ("q4_1", GGML_TYPE_Q4_1)]


def test_block_layouts():
    # lib/ggml.c:590-626
    assert (Q4_0_DT.itemsize, Q4_1_DT.itemsize, Q8_0_DT.itemsize) == (20, 24, 40)
    assert Q4_0_DT.fields["qs"][1] == 4 and Q4_1_DT.fields["qs"][1] == 8 and Q8_0_DT.fields["qs"][1] == 8


def test_q8_0_matches_golden(oracle, golden_rowfns):
    k, g = golden_rowfns
    assert np.array_equal(oracle.quantize_q8_0(g["x"]), g["q8"])


def test_q8_0_edge_semantics(oracle):
    """Round-half-even with id = 127/amax (lib/ggml.c:1362-1378), not roundf with 1/d."""
    x = np.zeros((1, 64), dtype=np.float32)
    x[0, :8] = [127.0, 0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.5]
    q = oracle.quantize_q8_0(x).view(Q8_0_DT)[0]
    assert q["d"][0] == np.float32(1.0)
    assert list(q["qs"][0][:8]) == [127, 0, 2, 2, 0, -2, -2, 4]
    assert q["s"][0] == np.float32(127 + 0 + 2 + 2 + 0 - 2 - 2 + 4)
    assert q["d"][1] == 0 and q["s"][1] == 0 and not q["qs"][1].any()      # all-zero block
    # the scalar fallback of the reference rounds half away from zero: different bytes
    qs = oracle.quantize_q8_0(x, scalar=True).view(Q8_0_DT)[0]
    assert list(qs["qs"][0][:8]) == [127, 1, 2, 3, -1, -2, -3, 4]


@pytest.mark.parametrize("name,t", TYPES)
def test_q4_quantize_dequantize_golden(oracle, golden_rowfns, name, t):
    k, g = golden_rowfns
    wq = oracle.quantize_q4(g["w"], t)
    assert np.array_equal(wq, g[f"{name}_w"])
    npq = (np_quantize_q4_0 if t == GGML_TYPE_Q4_0 else np_quantize_q4_1)(g["w"])
    assert np.array_equal(npq, g[f"{name}_w"])
    assert np.array_equal(oracle.dequantize_q4(wq, t, k).view(np.uint32), g[f"{name}_deq"].view(np.uint32))


@pytest.mark.parametrize("name,t", TYPES)
def test_mul_mat_golden_bit_exact(oracle, golden_rowfns, name, t):
    """The restated 8-lane fp32 accumulation reproduces the AVX2 build bit for bit."""
    k, g = golden_rowfns
    got = oracle.mul_mat_q(g[f"{name}_w"], g["x"], t)
    assert np.array_equal(got.view(np.uint32), g[f"{name}_mul_mat"].view(np.uint32))


@pytest.mark.parametrize("name,t", TYPES)
def test_exact_accumulation_bounds_reference(oracle, golden_rowfns, name, t):
    """fp32 reordering budget: the reference's own result sits within 2e-6 * sum|d q| of the
    order-free (double) value.  The GPU tests hold the CUDA kernels to the same budget."""
    k, g = golden_rowfns
    ex, mag = oracle.mul_mat_q_exact(g[f"{name}_w"], g["x"], t)
    err = np.abs(g[f"{name}_mul_mat"].astype(np.float64) - ex)
    assert np.all(err <= 2e-6 * mag + 1e-30)


def test_get_rows_is_dequantize(oracle, golden_rowfns):
    k, g = golden_rowfns
    ids = np.array([3, 0, 3, g["w"].shape[0] - 1], dtype=np.int32)
    for name, t in TYPES:
        rows = oracle.get_rows_q(g[f"{name}_w"], ids, t, k)
        assert np.array_equal(rows, g[f"{name}_deq"][ids])


# ---- live reference (oracle/_ref) on fresh inputs ---------------------------------------------
@pytest.mark.parametrize("k", [64, 4096, 11008])
def test_live_reference_rowfns(oracle, ref, k):
    rng = np.random.default_rng(k)
    x = (rng.standard_normal((16, k)) * rng.uniform(0.01, 30.0, (16, 1))).astype(np.float32)
    assert np.array_equal(oracle.quantize_q8_0(x), ref.quantize_q8_0(x))
    w = (rng.standard_normal((24, k)) * 0.02).astype(np.float32)
    for name, t in TYPES:
        wq = ref.quantize_q4_reference(w, t)
        assert np.array_equal(oracle.quantize_q4(w, t), wq)
        assert np.array_equal(oracle.dequantize_q4(wq, t, k), ref.dequantize_q4(wq, t, k))
        a, b = oracle.mul_mat_q(wq, x[:3], t), ref.mul_mat_q(wq, x[:3], t)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ---- the ops of attach_lora / detach_lora and the SIMD quantiser slot (row a2) ---------------------
GOLDEN_LORA = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden", "lora_ops.npz")


@pytest.fixture(scope="module")
def lora_golden():
    return np.load(GOLDEN_LORA)


@pytest.mark.parametrize("name,t", TYPES)
def test_simd_quantizers_match_golden(oracle, lora_golden, name, t):
    """quantize_fns[type].quantize_row_q = the AVX2 quantisers (lib/ggml.c:739-803, :965-1038): id = 7/amax and round-half-even,
    NOT the _reference arithmetic -- the fixture holds rows on which the two differ."""
    g = lora_golden
    got = oracle.quantize_q4_simd(g["w"], t)
    assert np.array_equal(got, g[f"{name}_simd"])
    assert not np.array_equal(got, g[f"{name}_base"]), "fixture must separate the SIMD and the _reference quantiser"


@pytest.mark.parametrize("r", [8, 16, 40, 64])
def test_f32_mul_mat_order_matches_golden(oracle, lora_golden, r):
    """B*A of a LoRA adapter: ggml_vec_dot_f32 in the AVX2 + FMA build's order (32-wide SIMD part, tree reduce, leftovers as the compiled loop adds them)."""
    g = lora_golden
    got = oracle.mul_mat_f32(g[f"A{r}"], g[f"B{r}"])
    assert np.array_equal(got.view(np.uint32), g[f"BA{r}"].view(np.uint32))


def test_f32_dot_leftovers_match_golden(oracle):
    """ggml_vec_dot_f32 for every leftover count n % 32 (tests/golden/f32_dot.npz, outputs of the reference library): the compiled
    reference adds leftovers in groups of 8 and one group of 4 as rounded product + rounded add and the last <= 3 as an fma.  The
    attention products of a prompt eval (inner length = number of positions) go through exactly this."""
    import os

    from oracle.gen_golden import F32_DOT_KS

    g = np.load(os.path.join(os.path.dirname(GOLDEN_LORA), "f32_dot.npz"))
    for k in F32_DOT_KS:
        got = oracle.mul_mat_f32(g[f"a{k}"], g[f"b{k}"])
        assert np.array_equal(got.view(np.uint32), g[f"out{k}"].view(np.uint32)), k


@pytest.mark.parametrize("name,t", TYPES)
@pytest.mark.parametrize("r", [8, 64])
def test_lora_merge_and_detach_match_golden(oracle, lora_golden, name, t, r):
    g = lora_golden
    ba = oracle.mul_mat_f32(g[f"A{r}"], g[f"B{r}"])
    merged = oracle.add_q_f32(g[f"{name}_base"], ba, t)
    assert np.array_equal(merged, g[f"{name}_merged{r}"])
    detached = oracle.add_q_f32(merged, -ba, t)
    assert np.array_equal(detached, g[f"{name}_detached{r}"])


@pytest.mark.parametrize("name,t", TYPES)
def test_simd_quantizers_match_live_reference(oracle, ref, name, t):
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((64, 512)) * rng.choice([1e-3, 1.0, 50.0], size=(64, 1))).astype(np.float32)
    x[0] = 0
    x[1, :32] = np.round(x[1, :32] * 2) / 2
    assert np.array_equal(oracle.quantize_q4_simd(x, t), ref.quantize_q4_simd(x, t))
