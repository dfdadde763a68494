import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "semilayer-wise-mixed-precision-quantization_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("SLQ_SYNTH_BATCHES", "2")
os.environ.setdefault("SLQ_SYNTH_BATCH", "4")
os.environ.setdefault("SLQ_SYNTH_HW", "64")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The CPU suite also needs libslq_b200.so (symbol tests) and the oracle .so."""
    import slq_build
    import slq_oracle
    slq_build.build()
    slq_oracle.build()
