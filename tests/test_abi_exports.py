"""CPU: the C-ABI libraries load without a GPU and export every function include/*.h declares; without a CUDA
device fl_init fails loudly and every compute entry point refuses to run (there is no CPU fallback in the product)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fastllama_b200", "lib")

DECL = re.compile(r"^\s*(?:const\s+)?(?:struct\s+\w+|unsigned\s+\w+|\w+)\s*\*{0,2}\s*\b((?:fl|ggml)_\w+)\s*\(", re.M)


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)           # comments mention function names too
    src = re.sub(r"//[^\n]*", "", src)
    names = set(DECL.findall(src))
    names -= {n for n in names if re.search(r"typedef[^;]*\(\s*\*\s*" + n, src)}     # function-pointer typedefs
    return sorted(names)


@pytest.mark.parametrize("header,lib", [("fl_cuda.h", "libfl_cuda.so"), ("fl_ggml.h", "libggml_b200.so")])
def test_every_declared_function_is_exported(header, lib):
    path = os.path.join(LIB, lib)
    if not os.path.exists(path):
        pytest.skip(f"{lib} not built (python -c 'import __graft_entry__ as g; g.build()')")
    if lib == "libggml_b200.so":
        C.CDLL(os.path.join(LIB, "libfl_cuda.so"), mode=C.RTLD_GLOBAL)
    so = C.CDLL(path)
    names = declared(header)
    assert len(names) > 30, names
    missing = [n for n in names if not hasattr(so, n)]
    assert not missing, f"{lib} does not export {missing}"


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    path = os.path.join(LIB, "libfl_cuda.so")
    if not os.path.exists(path):
        pytest.skip("libfl_cuda.so not built")
    so = C.CDLL(path)
    so.fl_last_error.restype = C.c_char_p
    assert so.fl_init(-1) != 0
    assert b"no CPU fallback" in so.fl_last_error() or b"CUDA" in so.fl_last_error()
    x = (C.c_float * 32)()
    y = (C.c_uint8 * 40)()
    assert so.fl_quantize_row_q8_0(x, y, 32) != 0          # refuses: not initialised
