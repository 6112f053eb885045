"""attach_lora / detach_lora ops and the SIMD quantiser slot of quantize_fns[] (row a2), against the
fixture tests/golden/lora_ops.npz = outputs of the reference LIBRARY (oracle/gen_golden.py: lora_ops): the graphs
    BA = mul_mat(loraA, loraB);  add_inplace(W_quantised, BA)            (attach, reference lib/llama.cpp:867-873, :907-913)
    add_inplace(W_quantised, scale(BA, -1))                              (detach, :929-941)
run through the ggml C API exactly like the reference's loader builds them.  All of it is bit-exact.

  * CPU: the reference library still reproduces the fixture (pins the fixture; needs oracle/_ref);
  * CPU: our host stack (quantised-add dispatch, in-place result in a persistent arena, leaf upload) on the CPU stand-in of the device layer;
  * GPU: the kernels of fastllama_b200/csrc/fl_lora_kernels.cu through libggml_b200 and through the C ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle.gen_golden import LORA_SHAPE
from oracle.pyoracle import REF_GGML_SO
from tests import ggml_api as G
from tests.mockbuild import ensure_mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "lora_ops.npz")
TYPES = [("q4_0", G.Q4_0), ("q4_1", G.Q4_1)]


def run_lora_graphs(lib_path, name, t, r):
    """-> (BA, merged bytes, detached bytes) computed by the library at lib_path."""
    gold = np.load(GOLDEN)
    g = G.Ggml(lib_path)
    k, m = LORA_SHAPE["k"], LORA_SHAPE["m"]
    # the weights live in their own (persistent) arena, the adapter and BA in the graph's arena -- as in the reference, where the
    # model context and the loader's scratch context are different buffers
    wa = g.context(4 << 20)
    tw = g.new_tensor_2d(wa.ctx, t, k, m)
    wa.set(tw, gold[f"{name}_base"])
    ar = g.context(64 << 20)
    ta = g.new_tensor_2d(ar.ctx, G.F32, r, k)
    ar.set(ta, gold[f"A{r}"])
    tb = g.new_tensor_2d(ar.ctx, G.F32, r, m)
    ar.set(tb, gold[f"B{r}"])
    ba = g.mul_mat(ar.ctx, ta, tb)
    res = g.add_inplace(ar.ctx, tw, ba)
    gf = G.new_graph()
    g.build_forward_expand(gf, res)
    g.graph_compute(ar.ctx, gf)
    sync = getattr(g.lib, "ggml_b200_sync_to_host", None)

    def weights():
        if sync is not None:                       # ours: the merged weights are on the device; fetch them for the comparison
            sync.argtypes, sync.restype = [G.C.c_void_p, G.C.c_size_t], None
            sync(tw.contents.data, gold[f"{name}_base"].nbytes)
        return wa.numpy(tw).reshape(m, -1).copy()

    if sync is not None:
        sync.argtypes, sync.restype = [G.C.c_void_p, G.C.c_size_t], None
        sync(ba.contents.data, k * m * 4)
    ba_v = ar.numpy(ba).reshape(m, k).copy()
    merged = weights()
    neg = g.scale(ar.ctx, ba, g.new_f32(ar.ctx, -1.0))
    res2 = g.add_inplace(ar.ctx, tw, neg)
    gf2 = G.new_graph()
    g.build_forward_expand(gf2, res2)
    g.graph_compute(ar.ctx, gf2)
    detached = weights()
    ar.free()
    wa.free()
    return ba_v, merged, detached


def check(outs, name, r):
    gold = np.load(GOLDEN)
    ba, merged, detached = outs
    assert np.array_equal(ba.view(np.uint32), gold[f"BA{r}"].view(np.uint32)), "B*A is not in the reference's summation order"
    assert np.array_equal(merged, gold[f"{name}_merged{r}"]), "merged weights differ from the reference's bytes"
    assert np.array_equal(detached, gold[f"{name}_detached{r}"])
    assert not np.array_equal(merged, gold[f"{name}_base"])


@pytest.mark.parametrize("name,t", TYPES)
def test_reference_library_reproduces_the_fixture(name, t):
    if not os.path.exists(REF_GGML_SO):
        pytest.skip("oracle/_ref not built")
    check(run_lora_graphs(REF_GGML_SO, name, t, 16), name, 16)


MOCK_RUN = r"""
import sys
sys.path.insert(0, sys.argv[1])
from tests.test_lora import check, run_lora_graphs
lib, name, t, r = sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
check(run_lora_graphs(lib, name, t, r), name, r)
print("OK")
"""


@pytest.mark.parametrize("name,t", TYPES)
@pytest.mark.parametrize("r", [8, 40])
def test_host_stack_on_cpu_mock(name, t, r):
    mock = ensure_mock()
    lib = os.path.join(mock, "libggml_b200.so")
    if not os.path.exists(lib):
        pytest.skip("mock not built")
    res = subprocess.run([sys.executable, "-c", MOCK_RUN, ROOT, lib, name, str(t), str(r)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("name,t", TYPES)
@pytest.mark.parametrize("r", [8, 16, 40, 64])
def test_lora_ops_on_gpu(name, t, r):
    from fastllama_b200.build import lib_path

    check(run_lora_graphs(lib_path("libggml_b200.so"), name, t, r), name, r)


@pytest.mark.gpu
@pytest.mark.parametrize("name,t", TYPES)
def test_quantize_fns_slot_is_the_simd_quantiser_on_gpu(name, t):
    """ggml_internal_get_quantize_fn(type).quantize_row_q on libggml_b200 (GPU-backed) = the reference's AVX2 quantiser, and the
    _reference slot still is the file quantiser (the fixture holds rows on which they differ)."""
    from fastllama_b200.build import lib_path
    from oracle.pyoracle import RefGgml

    gold = np.load(GOLDEN)
    ours = RefGgml(lib_path("libggml_b200.so"))          # same ctypes driver, our library
    assert np.array_equal(ours.quantize_q4_simd(gold["w"], t), gold[f"{name}_simd"])
    assert np.array_equal(ours.quantize_q4_reference(gold["w"], t), gold[f"{name}_base"])
    x = (np.random.default_rng(9).standard_normal((8, 4096)) * 3).astype(np.float32)
    assert np.array_equal(ours.quantize_q8_0(x), gold_q8(x))


def gold_q8(x):
    from oracle.pyoracle import Oracle

    return Oracle().quantize_q8_0(x)
