import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


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


@pytest.fixture(scope="session")
def golden_env():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "reversi_env.npz"))


@pytest.fixture(scope="session")
def golden_ttt():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "ttt_env.npz"))


@pytest.fixture(scope="session")
def golden_mcts():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "mcts.npz"))


@pytest.fixture(scope="session")
def golden_mcts_vl():
    """virtual-loss searches of the Python definition run on the live reference boards (make_golden.py --only-vl)"""
    import numpy as np

    return np.load(os.path.join(GOLDEN, "mcts_vl.npz"))
