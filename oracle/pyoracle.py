"""ctypes front-ends for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Two libraries live behind this module:

* ``libq4oracle.so``  -- our plain-C restatement (oracle/q4_oracle.c), always buildable.
* ``_ref/libggml_ref.so`` -- the reference's own lib/ggml.c compiled in place by
  oracle/Makefile (exists wherever /root/reference existed at build time; the built
  file travels to the GPU box).  Reached through the hook the reference exports for
  exactly this purpose, ``ggml_internal_get_quantize_fn`` (include/ggml.h:841-862,
  lib/ggml.c:1769-1773).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libq4oracle.so")
REF_GGML_SO = os.path.join(HERE, "_ref", "libggml_ref.so")
REF_PYFASTLLAMA_SO = os.path.join(HERE, "_ref", "pyfastllama_ref.so")

GGML_TYPE_F32, GGML_TYPE_F16, GGML_TYPE_Q4_0, GGML_TYPE_Q4_1, GGML_TYPE_Q8_0 = 0, 1, 2, 3, 6
BLOCK_BYTES = {GGML_TYPE_Q4_0: 20, GGML_TYPE_Q4_1: 24, GGML_TYPE_Q8_0: 40}
QK = 32


def build_oracle(force: bool = False) -> str:
    """Compile oracle/q4_oracle.c (and, where /root/reference exists, oracle/_ref)."""
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
        os.path.join(HERE, "q4_oracle.c")
    ):
        subprocess.check_call(["make", "-C", HERE, "all"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/lib") and not os.path.exists(REF_PYFASTLLAMA_SO):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


def have_ref() -> bool:
    return os.path.exists(REF_GGML_SO)


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Our C restatement (oracle/q4_oracle.c)."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build_oracle()
        self.lib = C.CDLL(path)
        L = self.lib
        for name in ("orc_quantize_row_q8_0", "orc_quantize_row_q8_0_scalar", "orc_quantize_row_q4_0", "orc_quantize_row_q4_1"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            getattr(L, name).restype = None
        for name in ("orc_dequantize_row_q4_0", "orc_dequantize_row_q4_1"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            getattr(L, name).restype = None
        for name in ("orc_vec_dot_q4_0_q8_0", "orc_vec_dot_q4_1_q8_0"):
            getattr(L, name).argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            getattr(L, name).restype = None
        L.orc_mul_mat_q_f32.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_mul_mat_q_f32.restype = C.c_int
        L.orc_mul_mat_q_exact.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_mul_mat_q_exact.restype = C.c_int
        L.orc_get_rows_q.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_rows_q.restype = C.c_int
        L.orc_quantize_q4.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_quantize_q4.restype = C.c_size_t
        for name in ("orc_quantize_row_q4_0_simd", "orc_quantize_row_q4_1_simd"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            getattr(L, name).restype = None
        L.orc_vec_dot_f32.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.orc_vec_dot_f32.restype = C.c_float
        L.orc_add_q_f32.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_add_q_f32.restype = C.c_int

    # --- row functions -------------------------------------------------------------
    def quantize_q8_0(self, x: np.ndarray, scalar: bool = False) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        k = x.shape[-1]
        rows = x.reshape(-1, k)
        out = np.empty((rows.shape[0], k // QK * 40), dtype=np.uint8)
        fn = self.lib.orc_quantize_row_q8_0_scalar if scalar else self.lib.orc_quantize_row_q8_0
        for r in range(rows.shape[0]):
            fn(_fptr(rows[r]), _fptr(out[r]), k)
        return out.reshape(x.shape[:-1] + (k // QK * 40,))

    def quantize_q4(self, x: np.ndarray, ggml_type: int) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        k = x.shape[-1]
        n = x.size
        out = np.empty(n // QK * BLOCK_BYTES[ggml_type], dtype=np.uint8)
        self.lib.orc_quantize_q4(ggml_type, _fptr(x), _fptr(out), n, k)
        return out.reshape(x.shape[:-1] + (k // QK * BLOCK_BYTES[ggml_type],))

    def quantize_q4_simd(self, x: np.ndarray, ggml_type: int) -> np.ndarray:
        """quantize_fns[type].quantize_row_q: the AVX2 quantisers (lib/ggml.c:739-803, :965-1038), not the _reference ones."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        k = x.shape[-1]
        rows = x.reshape(-1, k)
        bb = BLOCK_BYTES[ggml_type]
        out = np.empty((rows.shape[0], k // QK * bb), dtype=np.uint8)
        fn = self.lib.orc_quantize_row_q4_0_simd if ggml_type == GGML_TYPE_Q4_0 else self.lib.orc_quantize_row_q4_1_simd
        for r in range(rows.shape[0]):
            fn(_fptr(rows[r]), _fptr(out[r]), k)
        return out.reshape(x.shape[:-1] + (k // QK * bb,))

    def mul_mat_f32(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        """ggml_mul_mat of two f32 matrices: a [Ma, K], b [Mb, K] -> [Mb, Ma] with out[j][i] = vec_dot_f32(a[i], b[j])
        in the AVX2 build's summation order (lib/ggml.c:2295-2325)."""
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        out = np.empty((b.shape[0], a.shape[0]), dtype=np.float32)
        for j in range(b.shape[0]):
            for i in range(a.shape[0]):
                out[j, i] = self.lib.orc_vec_dot_f32(a.shape[1], _fptr(a[i]), _fptr(b[j]))
        return out

    def add_q_f32(self, w: np.ndarray, x: np.ndarray, ggml_type: int) -> np.ndarray:
        """ggml_compute_forward_add_q_f32 (lib/ggml.c:6414-6520): dequantise, add, re-quantise with the SIMD quantiser."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        rows, k = x.shape
        w = np.ascontiguousarray(w, dtype=np.uint8).reshape(rows, -1)
        out = np.empty_like(w)
        assert self.lib.orc_add_q_f32(ggml_type, rows, k, _fptr(w), _fptr(x), _fptr(out)) == 0
        return out

    def dequantize_q4(self, w: np.ndarray, ggml_type: int, k: int) -> np.ndarray:
        w = np.ascontiguousarray(w, dtype=np.uint8).reshape(-1, k // QK * BLOCK_BYTES[ggml_type])
        out = np.empty((w.shape[0], k), dtype=np.float32)
        fn = self.lib.orc_dequantize_row_q4_0 if ggml_type == GGML_TYPE_Q4_0 else self.lib.orc_dequantize_row_q4_1
        for r in range(w.shape[0]):
            fn(_fptr(w[r]), _fptr(out[r]), k)
        return out

    def vec_dot(self, wrow: np.ndarray, q8row: np.ndarray, ggml_type: int, k: int) -> np.float32:
        s = np.zeros(1, dtype=np.float32)
        fn = self.lib.orc_vec_dot_q4_0_q8_0 if ggml_type == GGML_TYPE_Q4_0 else self.lib.orc_vec_dot_q4_1_q8_0
        fn(k, _fptr(s), _fptr(np.ascontiguousarray(wrow)), _fptr(np.ascontiguousarray(q8row)))
        return s[0]

    # --- the op --------------------------------------------------------------------
    def mul_mat_q(self, w: np.ndarray, x: np.ndarray, ggml_type: int) -> np.ndarray:
        """w: [M, K/32*bb] uint8, x: [N, K] f32 -> dst [N, M] f32 (ggml: ne0=M contiguous)."""
        w = np.ascontiguousarray(w, dtype=np.uint8)
        x = np.ascontiguousarray(x, dtype=np.float32)
        m, (n, k) = w.shape[0], x.shape
        dst = np.empty((n, m), dtype=np.float32)
        rc = self.lib.orc_mul_mat_q_f32(ggml_type, m, k, n, _fptr(w), _fptr(x), _fptr(dst))
        assert rc == 0, rc
        return dst

    def mul_mat_q_exact(self, w: np.ndarray, x: np.ndarray, ggml_type: int):
        """Order-free (double) accumulation of the same block arithmetic, plus the magnitude
        sum_i |d_i q_i| that scales the admissible fp32 reordering error."""
        w = np.ascontiguousarray(w, dtype=np.uint8)
        x = np.ascontiguousarray(x, dtype=np.float32)
        m, (n, k) = w.shape[0], x.shape
        dst = np.empty((n, m), dtype=np.float64)
        mag = np.empty((n, m), dtype=np.float64)
        rc = self.lib.orc_mul_mat_q_exact(ggml_type, m, k, n, _fptr(w), _fptr(x), _fptr(dst), _fptr(mag))
        assert rc == 0, rc
        return dst, mag

    def get_rows_q(self, w: np.ndarray, ids: np.ndarray, ggml_type: int, k: int) -> np.ndarray:
        w = np.ascontiguousarray(w, dtype=np.uint8)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        dst = np.empty((ids.size, k), dtype=np.float32)
        rc = self.lib.orc_get_rows_q(ggml_type, k, ids.size, _fptr(w), _fptr(ids), _fptr(dst))
        assert rc == 0, rc
        return dst


class _QuantizeFns(C.Structure):
    # quantize_fns_t, include/ggml.h:850-862
    _fields_ = [
        ("dequantize_row_q", C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int)),
        ("quantize_row_q", C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int)),
        ("quantize_row_q_reference", C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int)),
        ("quantize_row_q_dot", C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int)),
        ("vec_dot_q", C.CFUNCTYPE(None, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)),
    ]


class RefGgml:
    """The reference's own row kernels, via ggml_internal_get_quantize_fn."""

    def __init__(self, path: str = REF_GGML_SO):
        self.lib = C.CDLL(path)
        self.lib.ggml_internal_get_quantize_fn.argtypes = [C.c_size_t]
        self.lib.ggml_internal_get_quantize_fn.restype = _QuantizeFns
        self.fns = {t: self.lib.ggml_internal_get_quantize_fn(t) for t in (GGML_TYPE_Q4_0, GGML_TYPE_Q4_1, GGML_TYPE_Q8_0)}

    def quantize_q8_0(self, x: np.ndarray) -> np.ndarray:
        """quantize_row_q_dot of a q4 type == quantize_row_q8_0 (lib/ggml.c:1735,1742)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        k = x.shape[-1]
        rows = x.reshape(-1, k)
        out = np.empty((rows.shape[0], k // QK * 40), dtype=np.uint8)
        fn = self.fns[GGML_TYPE_Q4_0].quantize_row_q_dot
        for r in range(rows.shape[0]):
            fn(_fptr(rows[r]), _fptr(out[r]), k)
        return out.reshape(x.shape[:-1] + (k // QK * 40,))

    def quantize_q4_reference(self, x: np.ndarray, ggml_type: int) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        k = x.shape[-1]
        rows = x.reshape(-1, k)
        bb = BLOCK_BYTES[ggml_type]
        out = np.empty((rows.shape[0], k // QK * bb), dtype=np.uint8)
        fn = self.fns[ggml_type].quantize_row_q_reference
        for r in range(rows.shape[0]):
            fn(_fptr(rows[r]), _fptr(out[r]), k)
        return out.reshape(x.shape[:-1] + (k // QK * bb,))

    def quantize_q4_simd(self, x: np.ndarray, ggml_type: int) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        k = x.shape[-1]
        rows = x.reshape(-1, k)
        bb = BLOCK_BYTES[ggml_type]
        out = np.empty((rows.shape[0], k // QK * bb), dtype=np.uint8)
        fn = self.fns[ggml_type].quantize_row_q
        for r in range(rows.shape[0]):
            fn(_fptr(rows[r]), _fptr(out[r]), k)
        return out.reshape(x.shape[:-1] + (k // QK * bb,))

    def dequantize_q4(self, w: np.ndarray, ggml_type: int, k: int) -> np.ndarray:
        w = np.ascontiguousarray(w, dtype=np.uint8).reshape(-1, k // QK * BLOCK_BYTES[ggml_type])
        out = np.empty((w.shape[0], k), dtype=np.float32)
        fn = self.fns[ggml_type].dequantize_row_q
        for r in range(w.shape[0]):
            fn(_fptr(w[r]), _fptr(out[r]), k)
        return out

    def vec_dot(self, wrow: np.ndarray, q8row: np.ndarray, ggml_type: int, k: int) -> np.float32:
        s = np.zeros(1, dtype=np.float32)
        self.fns[ggml_type].vec_dot_q(k, _fptr(s), _fptr(np.ascontiguousarray(wrow)), _fptr(np.ascontiguousarray(q8row)))
        return s[0]

    def mul_mat_q(self, w: np.ndarray, x: np.ndarray, ggml_type: int) -> np.ndarray:
        """The loop nest of ggml_compute_forward_mul_mat_q_f32 (lib/ggml.c:8105-8163) driven over
        the reference's own row kernels."""
        w = np.ascontiguousarray(w, dtype=np.uint8)
        x = np.ascontiguousarray(x, dtype=np.float32)
        m, (n, k) = w.shape[0], x.shape
        q8 = self.quantize_q8_0(x)
        dst = np.empty((n, m), dtype=np.float32)
        vd = self.fns[ggml_type].vec_dot_q
        s = np.zeros(1, dtype=np.float32)
        for j in range(n):
            for i in range(m):
                vd(k, _fptr(s), _fptr(w[i]), _fptr(q8[j]))
                dst[j, i] = s[0]
        return dst


# ---------------------------------------------------------------------------------------
# numpy views of the block formats (lib/ggml.c:590-626); used by tests and the GGJT writer
# ---------------------------------------------------------------------------------------
Q4_0_DT = np.dtype([("d", "<f4"), ("qs", "u1", (16,))])
Q4_1_DT = np.dtype([("d", "<f4"), ("m", "<f4"), ("qs", "u1", (16,))])
Q8_0_DT = np.dtype([("d", "<f4"), ("s", "<f4"), ("qs", "i1", (32,))])
assert Q4_0_DT.itemsize == 20 and Q4_1_DT.itemsize == 24 and Q8_0_DT.itemsize == 40


def _roundf(v: np.ndarray) -> np.ndarray:
    """C roundf (half away from zero), exact: v - trunc(v) is exact in fp32."""
    t = np.trunc(v)
    frac = v - t
    return t + np.where(np.abs(frac) >= np.float32(0.5), np.sign(v), np.float32(0.0)).astype(np.float32)


def np_quantize_q4_0(x: np.ndarray) -> np.ndarray:
    """Vectorised numpy restatement of quantize_row_q4_0_reference (lib/ggml.c:630-664) for
    writing large synthetic model files; checked bit-for-bit against the C oracle in tests."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    k = x.shape[-1]
    xb = x.reshape(-1, QK)
    amax = np.abs(xb).max(axis=1)
    d = (amax / np.float32(7.0)).astype(np.float32)
    with np.errstate(divide="ignore"):
        idv = np.where(d != 0, np.float32(1.0) / d, np.float32(0.0)).astype(np.float32)
    v = xb * idv[:, None]
    q = _roundf(v).astype(np.int8).astype(np.int16) + 8
    q = q.astype(np.uint8)
    out = np.empty(xb.shape[0], dtype=Q4_0_DT)
    out["d"] = d
    out["qs"] = q[:, 0::2] | (q[:, 1::2] << 4)
    return out.view(np.uint8).reshape(x.shape[:-1] + (k // QK * 20,))


def np_quantize_q4_1(x: np.ndarray) -> np.ndarray:
    """numpy restatement of quantize_row_q4_1_reference (lib/ggml.c:917-956)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    k = x.shape[-1]
    xb = x.reshape(-1, QK)
    mn = xb.min(axis=1)
    mx = xb.max(axis=1)
    d = ((mx - mn) / np.float32(15.0)).astype(np.float32)
    with np.errstate(divide="ignore"):
        idv = np.where(d != 0, np.float32(1.0) / d, np.float32(0.0)).astype(np.float32)
    v = (xb - mn[:, None]) * idv[:, None]
    q = _roundf(v).astype(np.uint8)
    out = np.empty(xb.shape[0], dtype=Q4_1_DT)
    out["d"] = d
    out["m"] = mn
    out["qs"] = q[:, 0::2] | (q[:, 1::2] << 4)
    return out.view(np.uint8).reshape(x.shape[:-1] + (k // QK * 24,))
