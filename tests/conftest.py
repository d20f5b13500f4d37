import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

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


def rel_l2(a, b) -> float:
    """||a - b|| / ||b|| in float64."""
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


@pytest.fixture(scope="session")
def ref_cases():
    return np.load(os.path.join(GOLDEN, "ref_cases.npz"))


@pytest.fixture(scope="session")
def ref_buffers():
    return np.load(os.path.join(GOLDEN, "ref_buffers.npz"))


@pytest.fixture(scope="session")
def ref_cfg1():
    return np.load(os.path.join(GOLDEN, "ref_cfg1.npz"))


def module_kwargs(case, dtype):
    return dict(im_size=case["im_size"], grid_size=case.get("grid_size"), numpoints=case.get("numpoints", 6),
                n_shift=case.get("n_shift"), table_oversamp=case.get("table_oversamp", 2 ** 10), dtype=dtype)
