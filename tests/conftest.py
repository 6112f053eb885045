import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle, build_oracle

    build_oracle()
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own lib/ggml.c (oracle/_ref), where it was built."""
    from oracle.pyoracle import RefGgml, build_oracle, have_ref

    build_oracle()
    if not have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    return RefGgml()


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session", params=[64, 256, 4096])
def golden_rowfns(request):
    import numpy as np

    return request.param, np.load(os.path.join(GOLDEN, f"rowfns_k{request.param}.npz"))
