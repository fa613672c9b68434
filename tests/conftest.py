"""pytest configuration: the `gpu` marker, KAT loading and shared helpers."""
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _int(x):
    return int(x, 0) if isinstance(x, str) else int(x)


@pytest.fixture(scope="session")
def kats():
    """Known-answer vectors transcribed from the reference's own tests (tests/golden)."""
    return json.loads((ROOT / "tests" / "golden" / "reference_kats.json").read_text())


@pytest.fixture(scope="session")
def to_int():
    return _int


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a box without a GPU: skip them loudly there.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (runs under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
