"""CPU: Model.save_state / load_state through the UNCHANGED reference bridge on top of libggml_b200, with the KV cache
living in "device" memory (the CPU stand-in of the device layer, tests/mock).  The bridge touches kv_self.{k,v}->data on
the host without telling ggml -- except for a ggml_nbytes(k) call right before each access (reference lib/llama.cpp:57-78),
which is the hook libggml_b200 uses to copy the cache back and to schedule its re-upload.  Checks: a state saved by us
resumes identically in us AND in the reference library, and a state saved by the reference resumes identically in us."""
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
from fastllama_b200.model import Model, QuietLogger
mode, lib, path, state, out = sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], sys.argv[6]
if "mock" in lib:
    C.CDLL(os.path.join(os.path.dirname(lib), "libfl_cuda.so"), mode=C.RTLD_GLOBAL)
m = Model(path, num_threads=2, n_ctx=64, n_batch=4, logger=QuietLogger(), library_path=lib)
toks = []
gen = lambda n: m.generate(lambda s: toks.append(s), num_tokens=n, temp=0.0, top_k=1, top_p=1.0, repeat_penalty=1.0)
if mode == "save":                       # prompt, 5 tokens, save, 8 more tokens, then load and redo the 8
    m.ingest("State files must survive a round trip.")
    gen(5)
    assert m.save_state(state)
    gen(8)
    first = list(toks[5:])
    assert m.load_state(state)
    del toks[:]
    gen(8)
    np.savez(out, first=np.array(first), again=np.array(toks), logits=m.get_logits_array())
else:                                    # resume from a state file written by somebody else
    assert m.load_state(state)
    gen(8)
    np.savez(out, again=np.array(toks), logits=m.get_logits_array())
m.close()
'''


def _run(tmp_path, mode, lib, model, state, tag):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = str(tmp_path / f"{tag}.npz")
    p = subprocess.run([sys.executable, str(script), ROOT, mode, lib, model, state, out], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert p.returncode == 0, p.stderr[-3000:]
    return np.load(out)


@pytest.mark.skipif(not os.path.exists(os.path.join(MOCK, "pyfastllama.so")), reason="tests/mock not built (needs the drop-in library)")
def test_state_round_trip_and_interop_with_reference(tmp_path):
    from fastllama_b200.ggjt import Q4_0, write_synthetic_numpy
    from oracle.pyoracle import REF_PYFASTLLAMA_SO, Oracle

    if not os.path.exists(REF_PYFASTLLAMA_SO):
        pytest.skip("oracle/_ref not built")
    orc = Oracle()
    model = str(tmp_path / "toy.bin")
    write_synthetic_numpy(model, Q4_0, n_vocab=512, n_embd=256, n_mult=256, n_head=4, n_layer=3, seed=7, std=0.01, quantize=lambda w, t: orc.quantize_q4(w, t))
    ours_lib = os.path.join(MOCK, "pyfastllama.so")

    ours = _run(tmp_path, "save", ours_lib, model, str(tmp_path / "ours.state"), "ours")
    assert list(ours["first"]) == list(ours["again"])                       # our own state resumes identically in us
    ref = _run(tmp_path, "save", REF_PYFASTLLAMA_SO, model, str(tmp_path / "ref.state"), "ref")
    assert list(ref["first"]) == list(ref["again"])
    assert os.path.getsize(tmp_path / "ours.state") == os.path.getsize(tmp_path / "ref.state")

    ref_from_ours = _run(tmp_path, "load", REF_PYFASTLLAMA_SO, model, str(tmp_path / "ours.state"), "ref_from_ours")
    assert list(ref_from_ours["again"]) == list(ours["first"])              # the KV cache we copied back is what the reference expects
    ours_from_ref = _run(tmp_path, "load", ours_lib, model, str(tmp_path / "ref.state"), "ours_from_ref")
    assert list(ours_from_ref["again"]) == list(ref["first"])               # a cache loaded on the host reaches the device
    tol = 2e-2 * np.abs(ref["logits"]).max()
    assert np.abs(ours_from_ref["logits"] - ref["logits"]).max() <= tol
