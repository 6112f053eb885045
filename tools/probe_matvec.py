"""GPU probe: time the decode matvec kernels at the LLaMA shapes (CUDA events on the library stream,
L2 flushed between launches) and print achieved GB/s of algorithmic bytes.  Not a bench line."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastllama_b200.cuda_abi import FlCuda  # noqa: E402

SHAPES = [("wq 7B", 2, 4096, 4096), ("w1 7B", 2, 11008, 4096), ("w2 7B", 2, 4096, 11008), ("out 7B", 2, 32000, 4096),
          ("wq 13B q4_1", 3, 5120, 5120), ("w1 13B q4_1", 3, 13824, 5120), ("w2 13B q4_1", 3, 5120, 13824),
          ("wq 65B", 2, 8192, 8192), ("w2 65B", 2, 8192, 22016)]


def main():
    fl = FlCuda()
    print(json.dumps(fl.device_props()))
    peak = 6480.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rng = np.random.default_rng(0)
    rows = []
    for name, t, m, k in SHAPES:
        bb = 20 if t == 2 else 24
        nb = k // 32
        w = rng.integers(0, 256, size=(m, nb, bb), dtype=np.uint8)
        w[:, :, 0:4] = np.frombuffer(np.float32(0.01).tobytes(), dtype=np.uint8)
        if t == 3:
            w[:, :, 4:8] = np.frombuffer(np.float32(-0.05).tobytes(), dtype=np.uint8)
        w = w.reshape(m, nb * bb)
        x = rng.standard_normal((1, k)).astype(np.float32)
        q8 = fl.quantize_q8_0(x)
        dW, dY, dD = fl.to_device(w), fl.to_device(q8), fl.alloc(m * 4)
        algo = m * nb * bb + nb * 40 + m * 4
        for impl in (1, 2):
            ms = C.c_float()
            fl.check(fl.lib.fl_dev_time_mul_mat_q(t, dW, nb * bb, m, k, dY, 1, dD, m, impl, 3, 256 << 20, C.byref(ms)))   # warm-up
            fl.check(fl.lib.fl_dev_time_mul_mat_q(t, dW, nb * bb, m, k, dY, 1, dD, m, impl, 20, 256 << 20, C.byref(ms)))
            gbs = algo / (ms.value * 1e-3) / 1e9
            rows.append((name, impl, ms.value * 1e3, gbs, gbs / peak))
            print(f"{name:14s} impl={impl} M={m:6d} K={k:6d}  {ms.value*1e3:8.2f} us  {gbs:8.1f} GB/s  {gbs/peak*100:5.1f}% of measured peak {peak:.0f}")
        for d in (dW, dY, dD):
            fl.free(d)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_matvec.json", "w") as f:
        json.dump(rows, f)


if __name__ == "__main__":
    main()
