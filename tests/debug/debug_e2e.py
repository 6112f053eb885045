import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fastllama_b200.build import lib_path
from fastllama_b200.ggjt import Q4_0, write_synthetic_numpy
from fastllama_b200.model import Model, QuietLogger, Logger
from oracle.pyoracle import REF_PYFASTLLAMA_SO, Oracle
orc = Oracle()
path = "/tmp/toy_dbg.bin"
write_synthetic_numpy(path, Q4_0, n_vocab=512, n_embd=256, n_mult=64, n_head=4, n_layer=3, seed=11, quantize=lambda w, t: orc.quantize_q4(w, t))
for lib in (REF_PYFASTLLAMA_SO, lib_path("pyfastllama.so")):
    m = Model(path, num_threads=8, n_ctx=128, n_batch=4, logger=QuietLogger(), library_path=lib)
    print("ingest", m.ingest("Hello world"), flush=True)
    toks = []
    print("gen", m.generate(lambda s: toks.append(s), num_tokens=3, temp=0.0, top_k=1, top_p=1.0, repeat_penalty=1.0), flush=True)
    lg = m.get_logits_array()
    print(os.path.basename(lib), toks, lg.shape, np.isnan(lg).sum(), lg[:6], flush=True)
    m.close()
