import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the library aborts the interpreter on its first call (no CPU fallback, by design): skip the
    gpu-marked tests instead of taking the CPU tests down with them when someone runs a plain `pytest tests` on a CPU box."""
    try:
        import torch
        has_gpu = torch.cuda.is_available() and torch.cuda.device_count() > 0
    except Exception:                                   # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (abinit_b200 has no CPU fallback; run with a B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """The product library, initialised on cuda:0. GPU tests call through this C-ABI only."""
    import abinit_b200
    abinit_b200.init(0)
    yield abinit_b200
    abinit_b200.finalize()
