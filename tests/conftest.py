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
