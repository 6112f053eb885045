#!/usr/bin/env python
"""bench.py -- tokens/sec of LLaMA-7B q4_0 greedy decode (n_batch = 1) on B200, BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (rank 0 only)

A "step" is one decoded token = one pass of the hot path (7*32+1 quantised matvecs, 4 129 423 360
algorithmic weight bytes) over a synthetic random-weight 7B q4_0 model (N(0, 0.02^2), seed 0, GGJT file
written once to $FASTLLAMA_BENCH_DIR or /tmp).  Keys of the JSON line:
  value      tokens/s with everything resident in HBM: K / (sum over the K timed evals of the CUDA-event
             time around the eval's kernels), whole job over all ranks (max over ranks of the time)
  e2e        tokens/s through the reference-facing API -- fastllama_b200.Model.generate() on the drop-in
             pyfastllama.so (the reference's unchanged bridge): wall clock around the call, which per step
             copies the token id host->device and the logits (+ embeddings row) device->host
  roofline   dominant kernel = k_decode_token, the persistent kernel that runs the whole decode step (one launch per
             token, reads every quantised weight once): algorithmic bytes per launch (4 129 423 360) / mean launch duration
             (CUDA events on the launching stream around the graph launch), against MEASURED_PEAKS.json's hbm_gbs;
             per_matrix_kernels = per-shape timings of the one-kernel-per-matrix-group path from extra instrumented steps
  cpu_baseline  the reference itself (oracle/_ref/pyfastllama_ref.so) on the host cores, bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROMPT = "The quick brown fox jumps over the lazy dog."
# weights-only bytes of the 7*n_layer+1 quantised matmuls
ALGO_BYTES_PER_TOKEN = {("7B", "q4_0"): 4129423360, ("7B", "q4_1"): 4955308032, ("13B", "q4_1"): 9638707200, ("65B", "q4_0"): 40638873600}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The reference's C++ layers print a banner and progress on the process's stdout (fd 1).  The driver wants exactly one JSON
# line there, so fd 1 is pointed at stderr for the whole run and the line is written to a private duplicate of the real stdout.
_REAL_STDOUT = None


def _claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


class _Stats(C.Structure):
    _fields_ = [("n_evals", C.c_uint64), ("last_eval_device_us", C.c_double), ("total_device_us", C.c_double),
                ("launches", C.c_uint64), ("graph_replays", C.c_uint64)]


class _KStat(C.Structure):
    _fields_ = [("type", C.c_int), ("M", C.c_int), ("K", C.c_int), ("N", C.c_int), ("launches", C.c_uint64),
                ("total_ms", C.c_double), ("algo_bytes_per_launch", C.c_double)]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": sorted(reasons),
                "samples": len(self.samples)}


def model_path(size: str, wtype_name: str) -> str:
    d = os.environ.get("FASTLLAMA_BENCH_DIR", "/tmp")
    return os.path.join(d, f"fastllama_b200_synth_{size}_{wtype_name}_seed0.bin")


def ensure_model(size: str, wtype: int, wtype_name: str) -> str:
    from fastllama_b200.ggjt import write_synthetic_gpu

    path = model_path(size, wtype_name)
    if not os.path.exists(path):
        t0 = time.time()
        tmp = path + f".tmp{os.getpid()}"
        n = write_synthetic_gpu(tmp, size=size, wtype=wtype, seed=0, std=0.02)
        os.replace(tmp, path)
        log(f"[bench] wrote synthetic {size} {wtype_name} model: {n/1e9:.2f} GB in {time.time()-t0:.1f}s -> {path}")
    return path


def run_reference(args, path, steps, threads=None):
    """The reference's own CPU implementation through its own C ABI (oracle/_ref/pyfastllama_ref.so)."""
    from fastllama_b200.model import Model, QuietLogger
    from oracle.pyoracle import REF_PYFASTLLAMA_SO

    if not os.path.exists(REF_PYFASTLLAMA_SO):
        raise RuntimeError("oracle/_ref/pyfastllama_ref.so is missing (build() must run where /root/reference exists)")
    ncpu = os.cpu_count() or 1
    threads = threads or min(ncpu, int(os.environ.get("FASTLLAMA_BENCH_CPU_THREADS", "32")))
    t0 = time.time()
    m = Model(path, num_threads=threads, n_ctx=512, n_batch=1, last_n_size=64, logger=QuietLogger(), library_path=REF_PYFASTLLAMA_SO)
    load_s = time.time() - t0
    m.ingest("Hi")                                   # 3 prompt tokens; the last one is evaluated by generate()
    stamps = []
    m.generate(lambda s: stamps.append(time.perf_counter()), num_tokens=args.warmup_cpu + steps, temp=0.0, top_k=1, top_p=1.0,
               repeat_penalty=1.0)
    m.close()
    stamps = stamps[args.warmup_cpu:]
    n = len(stamps) - 1
    tps = n / (stamps[-1] - stamps[0]) if n > 0 else 0.0
    return {"value": tps, "unit": "tokens/s", "cores": threads, "kind": "reference",
            "sample": f"{n} greedy decode tokens of the same synthetic 7B q4_0 file after {args.warmup_cpu} warm-up tokens, "
                      f"reference pyfastllama (AVX2 build) with num_threads={threads} of {ncpu} host cpus; model load {load_s:.1f}s not counted"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="7B")
    ap.add_argument("--wtype", default="q4_0", choices=["q4_0", "q4_1"])
    ap.add_argument("--cpu-steps", type=int, default=12, help="decode tokens of the CPU baseline sample")
    ap.add_argument("--warmup-cpu", type=int, default=2)
    ap.add_argument("--profile-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    _claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wtype = {"q4_0": 2, "q4_1": 3}[args.wtype]

    dist = None
    if world > 1 and args.impl == "ours":
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl")
        dist = dist_mod

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return                                                    # rank 0 alone runs and prints the reference arm
        path = model_path(args.size, args.wtype)
        if not os.path.exists(path):
            path = ensure_model(args.size, wtype, args.wtype)        # needs the GPU generator once
        steps = min(args.steps, int(os.environ.get("FASTLLAMA_BENCH_REF_MAX_STEPS", "24")))
        args.warmup_cpu = min(args.warmup, 3)
        t0 = time.perf_counter()
        cb = run_reference(args, path, steps)
        line = {"impl": "reference", "metric": f"tokens/sec LLaMA-{args.size} {args.wtype} decode (n_batch=1, greedy)", "value": cb["value"], "unit": "tokens/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup_cpu, "ms_per_step": 1000.0 / cb["value"] if cb["value"] else None,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"LLaMA-{args.size} {args.wtype} decode n_batch=1 n_ctx=512, reference CPU path", "cpu_threads": cb["cores"]},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wall_s": time.perf_counter() - t0}
        emit(line)
        return

    # ---------------------------------------------------------------- our arm
    os.environ.setdefault("FASTLLAMA_DEVICE", str(local_rank))
    from fastllama_b200.build import lib_path
    from fastllama_b200.cuda_abi import FlCuda
    from fastllama_b200.model import Model, QuietLogger

    fl = FlCuda()                                     # fails loudly without the CUDA library / a B200
    props = fl.device_props()
    tp = world > 1 and not os.environ.get("FASTLLAMA_BENCH_REPLICAS")
    if tp:
        # tensor parallelism: one NCCL communicator over all ranks; rank 0's unique id travels by torch.distributed
        import torch

        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = C.create_string_buffer(128)
            fl.check(fl.lib.fl_comm_unique_id(raw))
            idbuf = torch.tensor(list(raw.raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(idbuf, 0)
        fl.check(fl.lib.fl_comm_init(rank, world, bytes(idbuf.cpu().numpy().tobytes())))
    if rank == 0:
        path = ensure_model(args.size, wtype, args.wtype)
    if dist:
        dist.barrier()
    path = model_path(args.size, args.wtype)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        a2 = argparse.Namespace(**vars(args))
        a2.warmup_cpu = args.warmup_cpu
        cpu_baseline = run_reference(a2, path, args.cpu_steps)
        log(f"[bench] cpu_baseline: {cpu_baseline['value']:.2f} tokens/s on {cpu_baseline['cores']} threads")

    ggml = C.CDLL(lib_path("libggml_b200.so"))
    ggml.ggml_b200_get_stats.argtypes = [C.POINTER(_Stats)]
    ggml.ggml_b200_set_profile.argtypes = [C.c_int]
    ggml.ggml_b200_get_kernel_stats.argtypes = [C.POINTER(_KStat), C.c_int]
    ggml.ggml_b200_get_kernel_stats.restype = C.c_int

    def stats():
        s = _Stats()
        ggml.ggml_b200_get_stats(C.byref(s))
        return s

    t0 = time.time()
    m = Model(path, num_threads=1, n_ctx=512, n_batch=1, last_n_size=64, logger=QuietLogger(), library_path=lib_path("pyfastllama.so"))
    log(f"[bench] rank {rank}: model loaded in {time.time()-t0:.1f}s on {props['name']}")
    assert m.ingest(PROMPT)

    stamps, dev_us = [], []

    def on_token(_s):
        stamps.append(time.perf_counter())
        dev_us.append(stats().last_eval_device_us)

    # warm-up: evaluates the prompt (uploads the weights on the first eval) + W decode tokens
    m.generate(on_token, num_tokens=args.warmup, temp=0.0, top_k=1, top_p=1.0, repeat_penalty=1.0)
    stamps.clear()
    dev_us.clear()
    if dist:
        dist.barrier()
    fl.check(fl.lib.fl_sync())
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = fl.lib.fl_launch_count()
    t_begin = time.perf_counter()
    m.generate(on_token, num_tokens=args.steps, temp=0.0, top_k=1, top_p=1.0, repeat_penalty=1.0)
    fl.check(fl.lib.fl_sync())
    t_end = time.perf_counter()
    launches = fl.lib.fl_launch_count() - launches0
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    n_tok = len(stamps)
    if n_tok < args.steps:
        log(f"[bench] rank {rank}: generation stopped after {n_tok} of {args.steps} tokens (EOS); rates use the tokens produced")
    wall = t_end - t_begin
    device_s = sum(dev_us) * 1e-6

    decode_mode = int(ggml.ggml_b200_decode_mode())      # 2 = one persistent kernel per token, 1 = one kernel per matrix group
    # per-matrix view: a few more decode steps on the one-kernel-per-matrix-group path, every launch bracketed by CUDA events
    ggml.ggml_b200_set_profile(1)
    m.generate(lambda s: None, num_tokens=args.profile_steps, temp=0.0, top_k=1, top_p=1.0, repeat_penalty=1.0)
    ks = (_KStat * 64)()
    nk = ggml.ggml_b200_get_kernel_stats(ks, 64)
    ggml.ggml_b200_set_profile(0)
    m.close()

    if dist:
        import torch

        t = torch.tensor([wall, device_s, float(n_tok)], dtype=torch.float64, device="cuda")
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        wall, device_s = float(mx[0]), float(mx[1])
        total_tokens = float(n_tok) if tp else float(sm[2])        # tensor parallel: every rank decodes the SAME stream
    else:
        total_tokens = float(n_tok)
    if rank != 0:
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    decode_k = [k for k in ks[:nk] if k.N == 1]
    per_shape = [{"type": "q4_0" if k.type == 2 else "q4_1", "M": k.M, "K": k.K, "launches": int(k.launches), "us_per_launch": 1e3 * k.total_ms / k.launches,
                  "gbs": (k.algo_bytes_per_launch / (k.total_ms / k.launches * 1e-3) / 1e9) if k.total_ms > 0 else None} for k in decode_k if k.launches]
    algo = ALGO_BYTES_PER_TOKEN.get((args.size, args.wtype))
    value = total_tokens / device_s if device_s > 0 else 0.0
    e2e = total_tokens / wall if wall > 0 else 0.0
    if decode_mode == 2 and algo and n_tok:
        # dominant kernel = k_decode_token: ONE launch per token that reads every quantised weight once.  Its duration is the
        # CUDA-event bracket around the graph launch on the library stream (embedding-row dequant, a 4-byte memset and the
        # kernel itself; the first two are < 0.5 % of it), averaged over the timed steps.
        launch_s = device_s / n_tok
        achieved = algo / launch_s / 1e9
        roof_kernel = "k_decode_token (persistent per-token kernel: every quantised matvec + attention of the decode step, 1 launch per token)"
        launches_timed = int(n_tok)
        us_per_launch = launch_s * 1e6
    else:
        tot_ms = sum(k.total_ms for k in decode_k)
        tot_bytes = sum(k.algo_bytes_per_launch * k.launches for k in decode_k)
        launches_timed = int(sum(k.launches for k in decode_k))
        achieved = tot_bytes / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
        roof_kernel = "k_mv_fused (all quantised decode matvecs of the token step, one launch per matrix group)"
        us_per_launch = 1e3 * tot_ms / launches_timed if launches_timed else None
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch, like `achieved`)
    traffic = None
    if decode_mode == 2 and args.size == "7B" and args.wtype == "q4_0" and world == 1:
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_token_kernel.json")))
            traffic = int(cap["dram_bytes_read"]) + int(cap["dram_bytes_write"])
        except Exception:
            traffic = None
    from fastllama_b200.ggjt import LLAMA_SIZES

    n_embd_model = LLAMA_SIZES[args.size][0]
    line = {
        "metric": f"tokens/sec LLaMA-{args.size} {args.wtype} decode (n_batch=1, greedy)", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": n_tok,
        "warmup": args.warmup, "ms_per_step": 1000.0 * device_s / n_tok if n_tok else None, "higher_is_better": True,
        "scaling": "strong" if (world > 1 and tp) else "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"LLaMA-{args.size} {args.wtype} greedy decode, n_batch=1, n_ctx=512, synthetic random weights N(0,0.02^2) seed 0",
                   "parallelism": "1 GPU" if world == 1 else (
                       (f"tp{world}: wq/wk/wv/w1/w3/output row-split, wo/w2 K-split; the 2 reductions per layer are fused into the persistent token kernel "
                        "(partial sums pushed into peer-mapped buffers over NVLink, cross-GPU flag barrier), 1 NCCL all-gather of the logits, all in the CUDA graph"
                        if decode_mode == 2 else
                        f"tp{world}: wq/wk/wv/w1/w3/output row-split, wo/w2 K-split, 2 NCCL all-reduces of n_embd fp32 per layer + 1 logits all-gather, in the CUDA graph")
                       if tp else f"{world} independent replicas"),
                   "l2": f"inputs ({(algo or 0) / 1e9:.2f} GB of weights per token) are {(algo or 0) / 126e6:.0f}x larger than L2; no flush needed",
                   "algorithmic_bytes_per_token": algo, "device": props["name"]},
        "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 32000 * 4 + n_embd_model * 4,
                "api": "fastllama_b200.Model.generate -> pyfastllama.so (reference bridge, unchanged) -> libggml_b200 -> libfl_cuda"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                     "kernel": roof_kernel, "peak_source": peak_src, "launches_timed": launches_timed, "us_per_launch": us_per_launch,
                     "algorithmic_bytes_per_launch": algo if decode_mode == 2 else None,
                     "per_matrix_kernels": per_shape, "whole_token_gbs": (algo * value / 1e9) if algo else None},
        "clocks": sampler.summary(),
    }
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    emit(line)


if __name__ == "__main__":
    main()
