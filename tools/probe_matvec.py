"""GPU probe: time the decode matvec kernels at the LLaMA shapes.  Each shape gets enough copies of
its weight matrix to exceed L2 several times over; launches rotate through the copies back to back
(CUDA events around the batch), i.e. every launch streams from HBM like in a real decode step.
Prints achieved GB/s of ALGORITHMIC bytes.  Not a bench line.

  python tools/probe_matvec.py [--only SUBSTR] [--iters N] [--impl 1,2]
"""
import argparse
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--impl", default="1,2")
    ap.add_argument("--mode", default="1", help="0 eager launches, 1 CUDA-graph batch (default); the read-kernel calibration (mode 2) always runs")
    ap.add_argument("--footprint-mb", type=int, default=600)
    args = ap.parse_args()
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
        if args.only and args.only not in name:
            continue
        bb = 20 if t == 2 else 24
        nb = k // 32
        w = rng.integers(0, 256, size=(m, nb, bb), dtype=np.uint8)
        w[:, :, 0:4] = np.frombuffer(np.float32(0.01).tobytes(), dtype=np.uint8)
        if t == 3:
            w[:, :, 4:8] = np.frombuffer(np.float32(-0.05).tobytes(), dtype=np.uint8)
        w = w.reshape(m, nb * bb)
        wbytes = w.nbytes
        stride = (wbytes + 255) & ~255
        ncopies = max(1, -(-(args.footprint_mb << 20) // stride))
        x = rng.standard_normal((1, k)).astype(np.float32)
        q8 = fl.quantize_q8_0(x)
        dW = fl.alloc(stride * ncopies)
        fl.check(fl.lib.fl_h2d(dW, w.ctypes.data_as(C.c_void_p), wbytes))
        for c in range(1, ncopies):
            fl.check(fl.lib.fl_d2d(dW + c * stride, dW, wbytes))
        fl.check(fl.lib.fl_sync())
        dY, dD = fl.to_device(q8), fl.alloc(m * 4)
        algo = m * nb * bb + nb * 40 + m * 4
        for impl in [0x200] + [int(i) | (int(args.mode) << 8) for i in args.impl.split(",")]:
            ms = C.c_float()
            call = lambda it: fl.check(fl.lib.fl_dev_time_mul_mat_q_rot(t, dW, nb * bb, m, k, dY, 1, dD, m, impl, it, 0, stride, ncopies, C.byref(ms)))
            call(ncopies)          # warm-up
            call(args.iters)
            gbs = algo / (ms.value * 1e-3) / 1e9
            rows.append((name, impl, ms.value * 1e3, gbs, gbs / peak))
            print(f"{name:14s} impl={'read' if impl == 0x200 else impl & 0xFF} M={m:6d} K={k:6d} copies={ncopies:3d}  {ms.value*1e3:8.2f} us  {gbs:8.1f} GB/s  {gbs/peak*100:5.1f}% of measured peak {peak:.0f}")
        for d in (dW, dY, dD):
            fl.free(d)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_matvec.json", "w") as f:
        json.dump(rows, f)


if __name__ == "__main__":
    main()
