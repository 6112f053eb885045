"""GPU probe: prompt-ingest matmuls (N = 128) of LLaMA-7B q4_0, plain warp-per-row kernel (impl 1) vs the tensor-core kernel
(impl 3, mma.sync m16n8k32 u8 x s8).  Prints ms per matmul, TMAC/s and the implied tokens/s of a 128-token batch.

  python tools/probe_ingest.py [N]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastllama_b200.cuda_abi import FlCuda  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
fl = FlCuda()
rng = np.random.default_rng(0)
SHAPES = [("wq/wk/wv/wo", 4096, 4096, 4 * 32), ("w1/w3", 11008, 4096, 2 * 32), ("w2", 4096, 11008, 32), ("output", 32000, 4096, 1)]
total = {1: 0.0, 3: 0.0}
for name, m, k, count in SHAPES:
    nb = k // 32
    w = rng.integers(0, 256, size=(m, nb, 20), dtype=np.uint8)
    w[:, :, 0:4] = np.frombuffer(np.float32(0.01).tobytes(), dtype=np.uint8)
    w = w.reshape(m, nb * 20)
    q8 = fl.quantize_q8_0(rng.standard_normal((N, k)).astype(np.float32))
    dW, dY, dD = fl.to_device(w), fl.to_device(q8), fl.alloc(m * N * 4)
    for impl in (3, 1):
        iters = 20 if impl == 3 else 3
        ms = C.c_float()
        fl.check(fl.lib.fl_dev_time_mul_mat_q(2, dW, nb * 20, m, k, dY, N, dD, m, impl, iters, 0, C.byref(ms)))
        macs = m * k * N
        total[impl] += ms.value * count
        print(f"{name:12s} {m:6d} x {k:6d} x N={N}: impl {impl}: {ms.value:8.3f} ms  {macs / ms.value / 1e9:8.2f} TMAC/s")
    for d in (dW, dY, dD):
        fl.free(d)
for impl in (3, 1):
    print(f"impl {impl}: all quantised matmuls of a {N}-token 7B batch: {total[impl]:8.1f} ms -> {N / total[impl] * 1e3:8.0f} tokens/s (matmuls only)")
