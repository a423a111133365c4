import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU-marked tests are skipped (not failed) on a box without a CUDA device or without the built library."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    has_lib = (ROOT / "pantea_b200" / "libpantea_b200.so").exists()
    if has_gpu and has_lib:
        return
    why = "no CUDA device" if not has_gpu else "libpantea_b200.so not built"
    skip = pytest.mark.skip(reason=f"needs a B200: {why}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir() -> Path:
    return GOLDEN
