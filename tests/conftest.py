import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(group):
    """tests/golden/<group>.npz -> {case: {name: torch tensor / python scalar}}."""
    raw = np.load(os.path.join(GOLDEN, group + ".npz"), allow_pickle=False)
    cases = {}
    for key in raw.files:
        case, name = key.split("/", 1)
        arr = raw[key]
        if arr.ndim == 0:
            val = arr.item()
        else:
            val = torch.from_numpy(arr.copy())
        cases.setdefault(case, {})[name] = val
    return cases


@pytest.fixture(scope="session")
def golden_attention():
    return load_golden("attention")


@pytest.fixture(scope="session")
def golden_gcn():
    return load_golden("gcn")


@pytest.fixture(scope="session")
def golden_model():
    return load_golden("model")


@pytest.fixture(scope="session")
def golden_v2():
    return load_golden("v2")
