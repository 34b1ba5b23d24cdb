import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from tests import _oracle
    return _oracle.load()


@pytest.fixture(scope="session")
def gpu_prover_factory():
    """Creates ExecutionProver objects on cuda:0; fails loudly when the CUDA library cannot run."""
    import encrypt_zkvm_b200 as ezk
    if ezk.device_count() == 0:
        pytest.fail("no CUDA device visible: gpu-marked tests must run on the GPU box (no CPU fallback exists)")
    return ezk
