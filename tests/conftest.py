import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _poison_uninitialised():
    """GAITB200_TEST_POISON=1: every float tensor that torch.empty() hands out on a CUDA device is filled with NaN, so a
    kernel that reads a buffer nobody wrote turns its outputs into NaN instead of depending on what ran before."""
    import torch
    real_empty = torch.empty

    def empty(*a, **k):
        t = real_empty(*a, **k)
        if t.is_cuda and t.is_floating_point() and t.numel():
            t.fill_(float("nan"))
        return t

    torch.empty = empty


if os.environ.get("GAITB200_TEST_POISON") == "1":
    _poison_uninitialised()


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(GOLDEN / f"{name}.npz", allow_pickle=False))
    return load


@pytest.fixture(scope="session")
def smpl_data():
    from gaitb200 import synthetic
    return synthetic.make_smpl_data(seed=0, variant="sparse")


@pytest.fixture(scope="session")
def smpl_data_dense():
    from gaitb200 import synthetic
    return synthetic.make_smpl_data(seed=1, variant="dense")
