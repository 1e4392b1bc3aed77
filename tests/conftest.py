"""pytest configuration: the `gpu` marker and shared fixtures.

`python -m pytest tests -m "not gpu"` runs on a CPU-only box (oracle vs golden fixtures, host logic,
C-ABI symbol checks); `-m gpu` runs the parity tests through the C ABI on a B200.
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = Path(__file__).resolve().parent / "golden"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN_CASES = ["scene_config", "kat_overlap", "star16", "star64", "grid4"]


def load_golden(name: str):
    data = dict(np.load(GOLDEN / f"{name}.npz"))
    zones = json.loads((GOLDEN / f"{name}.zones.json").read_text())
    return data, zones


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    data, zones = load_golden(request.param)
    return request.param, data, zones


@pytest.fixture(scope="session")
def built_lib():
    """Build (or reuse) libopd_b200.so; nvcc cross-compiles without a GPU."""
    from office_person_detection_vit_b200.build import build

    return build()
