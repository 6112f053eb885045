"""GPU, last in the run: the persistent token kernel at LLaMA-65B q4_0 matrix shapes (n_embd 8192, 64 heads, n_ff 22016; one layer
and a small LM head) against the CPU model of the same steps -- the reference's bits at the shapes of BASELINE.json's fifth
configuration.  (The 65B model itself, 40.6 GB, is not generated in the test suite.)"""
import pytest

from oracle.pyoracle import GGML_TYPE_Q4_0
from tests.test_gpu_fused import cpu_model, fl, test_token_kernel_has_the_reference_bits as _check  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


def test_token_kernel_has_the_reference_bits_at_65b_shapes(fl, cpu_model):  # noqa: F811
    _check.__wrapped__(fl, cpu_model, GGML_TYPE_Q4_0, 8192, 64, 22016, 2000, 512, 130, 1) if hasattr(_check, "__wrapped__") else \
        _check(fl, cpu_model, GGML_TYPE_Q4_0, 8192, 64, 22016, 2000, 512, 130, 1)
