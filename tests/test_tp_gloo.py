"""CPU, world_size 2 over gloo: the tensor-parallel decode plan on the CPU stand-in of the device
layer (tests/mock).  Each rank runs the unchanged reference bridge over libggml_b200; the mock delegates the two
collectives to torch.distributed and maps POSIX shared memory between the rank processes where the GPUs map peer HBM.  Checks: both
ranks produce the single-process token sequence AND THE SAME LOGIT BITS: every matrix is row-split (wq/wk/wv by heads, w1/w3, wo, w2 and
the output matrix by rows), a row is summed over the whole K on one rank, and the activation vectors are gathered element by element as
dataflow (LL) vectors whose epoch polling is real here (the producer is the other process)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from tests.mockbuild import ensure_mock  # noqa: E402

MOCK = ensure_mock()

WORKER = r'''
import ctypes as C, os, sys, numpy as np
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from fastllama_b200.model import Model, QuietLogger
rank, world, mock, path, out = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), sys.argv[2], sys.argv[3], sys.argv[4]
if world > 1:
    dist.init_process_group("gloo", rank=rank, world_size=world)
lib = C.CDLL(os.path.join(mock, "libfl_cuda.so"), mode=C.RTLD_GLOBAL)
CB = C.CFUNCTYPE(None, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)
def coll(kind, send, recv, n):
    if kind == 0:
        a = np.ctypeslib.as_array((C.c_float * n).from_address(send))
        t = torch.from_numpy(a); dist.all_reduce(t)
    else:
        s = torch.from_numpy(np.ctypeslib.as_array((C.c_float * n).from_address(send)).copy())
        r = np.ctypeslib.as_array((C.c_float * (n * world)).from_address(recv))
        parts = [torch.empty(n) for _ in range(world)]
        dist.all_gather(parts, s)
        r[:] = torch.cat(parts).numpy()
cb = CB(coll)
lib.fl_mock_set_collective(cb, rank, world)
scenario = sys.argv[6] if len(sys.argv) > 6 else "decode"
m = Model(path, num_threads=2, n_ctx=64 if scenario == "decode" else 128, n_batch=int(sys.argv[5]), logger=QuietLogger(), library_path=os.path.join(mock, "pyfastllama.so"))
m.ingest("Tensor parallel decode over two ranks.")
toks = []
gen = lambda n: m.generate(lambda s: toks.append(s), num_tokens=n, temp=0.0, top_k=1, top_p=1.0, repeat_penalty=1.0)
if scenario == "reingest":
    # sharded decode steps, then a replicated multi-token eval that reads the whole KV cache (needs the KV gather), then
    # a state file written from a gathered cache and resumed
    gen(5)
    m.ingest(" And a second prompt that attends to all of it.")
    gen(4)
    assert m.save_state(out + f".rank{rank}.state")
    gen(3)
    first = list(toks[-3:])
    assert m.load_state(out + f".rank{rank}.state")
    gen(3)
    assert list(toks[-3:]) == first, (toks[-3:], first)
else:
    gen(10)
mode = C.CDLL(os.path.join(mock, "libggml_b200.so")).ggml_b200_decode_mode()
np.savez(out + f".rank{rank}.npz", toks=np.array(toks), logits=m.get_logits_array(), mode=mode)
m.close()
'''


@pytest.mark.skipif(not os.path.exists(os.path.join(MOCK, "pyfastllama.so")), reason="tests/mock not built (needs the drop-in library)")
@pytest.mark.parametrize("n_batch,world", [(1, 2), (8, 2), (1, 4)])
def test_tensor_parallel_decode_matches_single_rank(tmp_path, n_batch, world, scenario="decode"):
    """The mock maps POSIX shared memory between the ranks like fl_comm_shared_alloc maps peer HBM, so the sharded plan runs as the
    token program: every step stores its row slice of a vector into both ranks' copies and the consumers poll the epochs."""
    peer = True
    from fastllama_b200.ggjt import Q4_0, write_synthetic_numpy
    from oracle.pyoracle import Oracle

    orc = Oracle()
    path = str(tmp_path / "toy.bin")
    # world 4 needs n_ff % (32 * 4) == 0: n_mult 256 gives n_ff 768
    write_synthetic_numpy(path, Q4_0, n_vocab=512, n_embd=256, n_mult=64 if world == 2 else 256, n_head=4, n_layer=3, seed=5, std=0.01,
                          quantize=lambda w, t: orc.quantize_q4(w, t))
    script = tmp_path / "worker.py"
    script.write_text(WORKER)

    def launch(world, tag):
        procs = []
        for r in range(world):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", OMP_NUM_THREADS="2")
            env.pop("FL_MOCK_SESSION", None)
            if peer:
                env["FL_MOCK_SESSION"] = f"{os.getpid()}_{tag}_{n_batch}"
            procs.append(subprocess.Popen([sys.executable, str(script), ROOT, MOCK, path, str(tmp_path / tag), str(n_batch), scenario], env=env,
                                          stdout=subprocess.DEVNULL, stderr=subprocess.PIPE))
        for p in procs:
            _, err = p.communicate(timeout=240)
            assert p.returncode == 0, err.decode()[-2000:]
        return [np.load(str(tmp_path / tag) + f".rank{r}.npz") for r in range(world)]

    single = launch(1, "w1")[0]
    tp = launch(world, f"w{world}")
    assert int(single["mode"]) == 2                                   # one rank: the token program
    for r in tp:
        assert int(r["mode"]) == 2                                    # two ranks: the token program as well
        assert list(r["toks"]) == list(single["toks"])
        assert np.array_equal(r["logits"].view(np.uint32), single["logits"].view(np.uint32))      # row-split + gather: the one-rank bits


@pytest.mark.skipif(not os.path.exists(os.path.join(MOCK, "pyfastllama.so")), reason="tests/mock not built (needs the drop-in library)")
def test_tensor_parallel_kv_gather_before_replicated_eval_and_state(tmp_path):
    """Decode steps shard the KV cache by head; a later multi-token eval and save_state need all heads: the ranks
    all-gather the sharded positions first (ggml_b200.cpp tp_gather_kv).  Same tokens as one rank, and each rank's state
    file resumes identically."""
    test_tensor_parallel_decode_matches_single_rank(tmp_path, 4, 2, scenario="reingest")
