import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def emulated_abi(monkeypatch):
    """Route lib.call to the torch restatement of the C-ABI contract (tests/abi_emulator.py) so the host
    logic can run on CPU tensors.  Test-only: the product has no such path."""
    import idccrn_b200
    import abi_emulator
    from idccrn_b200 import lib

    monkeypatch.setattr(lib, "call", abi_emulator.call)
    monkeypatch.setattr(lib, "require_f32_cuda", lambda t, what: t.contiguous())
    return abi_emulator


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return {k: v for k, v in np.load(os.path.join(GOLDEN, name + ".npz")).items()}
    return load


@pytest.fixture(params=["tc", "simt"])
def gemm_mode(request):
    """Run a test once per tap-GEMM implementation: tcgen05 on split-bf16 planes / fp32 SIMT on fp32 planes."""
    import idccrn_b200
    from idccrn_b200 import ops
    old = ops.GEMM_MODE[0]
    ops.set_gemm_mode(request.param)
    yield request.param
    ops.set_gemm_mode(old)
