"""pytest configuration: markers, import paths, shared fixtures.

`-m "not gpu"` : oracle vs golden vectors, host logic, ABI exports, gloo world-size-2 (CPU only).
`-m gpu`       : parity tests proper -- the CUDA path through the C ABI against the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cuda-spmv-benchmark_b200", "python"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import orc as _orc
    _orc.lib()  # builds liboracle.so on first use
    return _orc


@pytest.fixture(scope="session")
def B():
    import spmv_b200
    if not os.path.exists(spmv_b200.LIB_PATH):
        spmv_b200.build()
    spmv_b200.load()
    return spmv_b200


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    torch.cuda.set_device(0)
    return torch
