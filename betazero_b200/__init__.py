"""betazero_b200 -- B200-native (sm_100a) self-play hot path for whaiproject/BetaZero.

Lockstep bitboard environments (Reversi, tic-tac-toe), batched warp-per-tree MCTS over SoA node
pools in HBM, and the leaf gather that feeds a PyTorch policy/value net, behind the C ABI of
``include/betazero_b200.h``.  The Python layer mirrors the reference's duck-typed board / player
interfaces.  There is no CPU fallback: every computation goes through
``libbetazero_b200.so``.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"


def build_library(force: bool = False) -> str:
    """Compile the CUDA library in-tree (nvcc, sm_100a)."""
    from . import build as _b

    return _b.build(force=force)
