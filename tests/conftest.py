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


@pytest.fixture(scope="module", autouse=True)
def _device_bounds_checks():
    """With a library built with -DBZ_BOUNDS_CHECK (BETAZERO_B200_LIB=build/exp/lib_bounds.so, profiles/sanitize.sh) every
    test module ends by reading the device-side violation records: the GPU suite doubles as the out-of-bounds check that
    compute-sanitizer would otherwise provide.  A release library has no such symbols and this is a no-op."""
    yield
    if "BETAZERO_B200_LIB" not in os.environ:
        return
    try:
        import ctypes

        import torch

        from betazero_b200 import _lib

        if not torch.cuda.is_available() or _lib._lib is None:
            return
        torch.cuda.synchronize()
        for name in ("bz_debug_checks_mcts", "bz_debug_checks_selfplay"):
            if hasattr(_lib._lib, name):
                rec = (ctypes.c_int * 4)()
                assert getattr(_lib._lib, name)(rec) == 0
                assert rec[3] == 0, f"{name}: check {rec[0]} failed in block {rec[1]}, thread {rec[2]} ({rec[3]} violations)"
    except ImportError:
        pass
