"""Builds the CPU stand-in of the device layer (tests/mock) on demand, so the host-logic tests do not depend on somebody
having run `make -C tests/mock` after the last build of the product libraries."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "mock", "build")


def ensure_mock() -> str:
    lib = os.path.join(ROOT, "fastllama_b200", "lib")
    if os.path.exists(os.path.join(lib, "libggml_b200.so")):
        try:
            subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "mock")], check=True, capture_output=True, timeout=300)
        except Exception:
            pass                      # the tests skip when the mock is missing
    return MOCK
