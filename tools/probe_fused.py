"""GPU probe: graph-batched timing of fl_dev_mv_fused at the LLaMA-7B decode shapes, rotating through
enough weight copies to defeat L2.  Compares prologue/epilogue variants.  Not a bench line."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastllama_b200.cuda_abi import FlCuda, FlMvArgs  # noqa: E402

fl = FlCuda()
rng = np.random.default_rng(0)
PEAK = 6480.8


def run(name, k, rows, pro, epi, iters=120):
    t, bb = 2, 20
    nb = k // 32
    mtot = sum(rows)
    wbytes = mtot * nb * bb
    w = rng.integers(0, 256, size=(mtot, nb, bb), dtype=np.uint8)
    w[:, :, 0:4] = np.frombuffer(np.float32(0.01).tobytes(), dtype=np.uint8)
    stride = (wbytes + 255) & ~255
    ncopies = max(2, -(-(500 << 20) // stride))
    dW = fl.alloc(stride * ncopies)
    fl.check(fl.lib.fl_h2d(dW, w.ctypes.data_as(C.c_void_p), wbytes))
    for c in range(1, ncopies):
        fl.check(fl.lib.fl_d2d(dW + c * stride, dW, wbytes))
    x = fl.to_device(rng.standard_normal(k).astype(np.float32))
    g = fl.to_device(np.ones(k, dtype=np.float32))
    b = fl.to_device(rng.standard_normal(k).astype(np.float32))
    res = fl.to_device(rng.standard_normal(max(rows)).astype(np.float32))
    outs = [fl.alloc(m * 4) for m in rows]
    n_ctx, n_embd, hd = 512, 4096, 128
    kc, vc = fl.alloc(n_ctx * n_embd * 4), fl.alloc(n_ctx * n_embd * 4)
    dnp = fl.to_device(np.array([100], dtype=np.int32))
    fl.check(fl.lib.fl_dev_rope_table(hd, n_ctx))
    fl.check(fl.lib.fl_sync())

    def args(copy):
        a = FlMvArgs()
        a.type, a.K, a.nseg, a.pro, a.epi = t, k, len(rows), pro, epi
        off = dW + copy * stride
        for i, m in enumerate(rows):
            a.seg_w[i], a.seg_rows[i], a.seg_dst[i] = off, m, outs[i]
            off += m * nb * bb
        a.x, a.gamma, a.b, a.res = x, g, b, res
        a.n_past, a.n_ctx, a.n_embd, a.head_dim, a.kcache, a.vcache = dnp, n_ctx, n_embd, hd, kc, vc
        return a

    keep = [args(i % ncopies) for i in range(iters)]
    fl.check(fl.lib.fl_dev_mv_fused(C.byref(keep[0])))          # attributes outside capture
    fl.check(fl.lib.fl_graph_begin_capture())
    for a in keep:
        fl.check(fl.lib.fl_dev_mv_fused(C.byref(a)))
    ge = C.c_void_p()
    fl.check(fl.lib.fl_graph_end_capture(C.byref(ge)))
    e0, e1 = fl.lib.fl_event_create(), fl.lib.fl_event_create()
    fl.check(fl.lib.fl_graph_launch(ge))
    fl.check(fl.lib.fl_event_record(e0))
    fl.check(fl.lib.fl_graph_launch(ge))
    fl.check(fl.lib.fl_event_record(e1))
    fl.check(fl.lib.fl_event_sync(e1))
    ms = C.c_float()
    fl.check(fl.lib.fl_event_elapsed_ms(e0, e1, C.byref(ms)))
    us = ms.value * 1e3 / iters
    gbs = wbytes / (us * 1e-6) / 1e9
    print(f"{name:28s} K={k:6d} M={mtot:6d} pro={pro} epi={epi}  {us:7.2f} us  {gbs:7.1f} GB/s  {gbs/PEAK*100:5.1f}%", flush=True)
    fl.check(fl.lib.fl_graph_destroy(ge))
    for d in [dW, x, g, b, res, kc, vc, dnp] + outs:
        fl.free(d)


only = sys.argv[1] if len(sys.argv) > 1 else ""
CASES = [("qkv plain/store", 4096, (4096, 4096, 4096), 0, 0), ("qkv rmsnorm/store", 4096, (4096, 4096, 4096), 1, 0), ("qkv rmsnorm/qkv", 4096, (4096, 4096, 4096), 1, 2),
         ("wo plain/store", 4096, (4096,), 0, 0), ("wo plain/resadd", 4096, (4096,), 0, 1),
         ("w13 plain/store", 4096, (11008, 11008), 0, 0), ("w13 rmsnorm/store", 4096, (11008, 11008), 1, 0),
         ("w2 plain/store", 11008, (4096,), 0, 0), ("w2 silumul/resadd", 11008, (4096,), 2, 1),
         ("head rmsnorm/store", 4096, (32000,), 1, 0)]
for c in CASES:
    if only in c[0]:
        run(*c)
