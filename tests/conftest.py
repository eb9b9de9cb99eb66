import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def O():
    """CPU oracle (test infrastructure)."""
    return ge.load_oracle()


@pytest.fixture(scope="session")
def cm():
    """the product binding; building happens in __graft_entry__.build()."""
    if not os.path.exists(os.path.join(ge.PKG_DIR, "libcudamat_b200.so")):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def pin():
    return np.load(os.path.join(GOLDEN, "ref_loader_csr.npz"))


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch
