import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "visual-rag-toolkit_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_index():
    with open(os.path.join(GOLDEN, "golden_index.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def pooling_golden():
    return np.load(os.path.join(GOLDEN, "pooling_golden.npz"))


@pytest.fixture(scope="session")
def maxsim_golden():
    return np.load(os.path.join(GOLDEN, "maxsim_golden.npz"))


@pytest.fixture(scope="session")
def retrieval_golden():
    return np.load(os.path.join(GOLDEN, "retrieval_golden.npz"))


def gpu_available() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
