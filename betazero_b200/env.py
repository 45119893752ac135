"""Batched lockstep environments on the GPU (kernels K1-K4, K6).

Functional API over SoA device tensors; each function is one launch of a hand-written sm_100a
kernel through the C ABI (include/betazero_b200.h).  Reversi boards are mover-relative
``(me, opp)`` pairs stored as ``torch.int64`` bit patterns of the uint64 bitboards
(bit = row*8 + col); actions are ``uint8`` (row*8 + col, 64 = pass).

Reference computations replaced (whaiproject/BetaZero, paths relative to its root):
  legal_mask      ReversiBoard.generate_possible_moves / is_valid_move  reversi_board.py:25-41,87-88
  apply           ReversiBoard.make_move + the pass of the driver       reversi_board.py:43-59, reversi_terminal.py:31-35
  terminal        ReversiBoard.is_game_over + get_score                 reversi_board.py:61-76
  planes          canonical form ``symbol * board``                     players.py:85
  ttt_*           TicTacToeBoard                                        tic_tac_toe_board.py:20-43
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

PASS = 64
N_ACTIONS = 65
TTT_ACTIONS = 9


# ------------------------------------------------------------------------------- host <-> device helpers
def to_device_u64(a, device="cuda") -> torch.Tensor:
    """uint64 host array (numpy / list of ints) -> int64 CUDA tensor with the same bits."""
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.uint64))
    return torch.from_numpy(arr.view(np.int64)).to(device, non_blocking=False)


def to_host_u64(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().view(np.uint64)


def to_device_u16(a, device="cuda") -> torch.Tensor:
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.uint16))
    return torch.from_numpy(arr.view(np.int16)).to(device)


def to_host_u16(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().view(np.uint16)


def _chk_boards(me: torch.Tensor, opp: torch.Tensor):
    if me.dtype != torch.int64 or opp.dtype != torch.int64:
        raise TypeError("bitboards must be torch.int64 tensors (bit patterns of uint64)")
    if me.shape != opp.shape or me.dim() != 1:
        raise ValueError("me/opp must be 1-D tensors of equal length")
    return me.numel()


# ------------------------------------------------------------------------------- Reversi
def reversi_init(n: int, size: int = 8, device="cuda"):
    """Start positions + player to move (+1 = X first, reversi_terminal.py:14)."""
    me = torch.empty(n, dtype=torch.int64, device=device)
    opp = torch.empty_like(me)
    player = torch.empty(n, dtype=torch.int8, device=device)
    L = _lib.load()
    _lib.check(L.bz_reversi_init(_lib.dptr(me), _lib.dptr(opp), _lib.dptr(player), n, size, _lib.stream_ptr()),
               "bz_reversi_init")
    return me, opp, player


def legal_mask(me: torch.Tensor, opp: torch.Tensor, size: int = 8, out: torch.Tensor | None = None) -> torch.Tensor:
    n = _chk_boards(me, opp)
    mask = torch.empty_like(me) if out is None else out
    L = _lib.load()
    _lib.check(L.bz_reversi_legal_mask(_lib.dptr(me), _lib.dptr(opp), _lib.dptr(mask), n, size, _lib.stream_ptr()),
               "bz_reversi_legal_mask")
    return mask


def apply(me: torch.Tensor, opp: torch.Tensor, action: torch.Tensor, size: int = 8, out=None):
    """Returns (me', opp', err) for the next mover; err[i] = 1 <=> ValueError("Invalid move")."""
    n = _chk_boards(me, opp)
    if action.dtype != torch.uint8 or action.numel() != n:
        raise TypeError("action must be a uint8 tensor of the same length")
    if out is None:
        me_o, opp_o = torch.empty_like(me), torch.empty_like(opp)
        err = torch.empty(n, dtype=torch.uint8, device=me.device)
    else:
        me_o, opp_o, err = out
    L = _lib.load()
    _lib.check(L.bz_reversi_apply(_lib.dptr(me), _lib.dptr(opp), _lib.dptr(action), _lib.dptr(me_o), _lib.dptr(opp_o),
                                  _lib.dptr(err), n, size, _lib.stream_ptr()), "bz_reversi_apply")
    return me_o, opp_o, err


def terminal(me: torch.Tensor, opp: torch.Tensor, size: int = 8):
    """Returns (over u8, winner_for_mover i8, cnt_me u8, cnt_opp u8)."""
    n = _chk_boards(me, opp)
    dev = me.device
    over = torch.empty(n, dtype=torch.uint8, device=dev)
    win = torch.empty(n, dtype=torch.int8, device=dev)
    cm = torch.empty(n, dtype=torch.uint8, device=dev)
    co = torch.empty(n, dtype=torch.uint8, device=dev)
    L = _lib.load()
    _lib.check(L.bz_reversi_terminal(_lib.dptr(me), _lib.dptr(opp), _lib.dptr(over), _lib.dptr(win), _lib.dptr(cm),
                                     _lib.dptr(co), n, size, _lib.stream_ptr()), "bz_reversi_terminal")
    return over, win, cm, co


def step_first_legal(me: torch.Tensor, opp: torch.Tensor, size: int = 8, out=None):
    """Fused K1+K2: (mask, action, me', opp') with action = lowest legal cell, or pass."""
    n = _chk_boards(me, opp)
    if out is None:
        mask = torch.empty_like(me)
        act = torch.empty(n, dtype=torch.uint8, device=me.device)
        me_o, opp_o = torch.empty_like(me), torch.empty_like(opp)
    else:
        mask, act, me_o, opp_o = out
    L = _lib.load()
    _lib.check(L.bz_reversi_step_first_legal(_lib.dptr(me), _lib.dptr(opp), _lib.dptr(mask), _lib.dptr(act),
                                             _lib.dptr(me_o), _lib.dptr(opp_o), n, size, _lib.stream_ptr()),
               "bz_reversi_step_first_legal")
    return mask, act, me_o, opp_o


def planes(me: torch.Tensor, opp: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Canonical bf16 network input [n, 2, 8, 8] (plane 0 = mover, plane 1 = opponent)."""
    n = _chk_boards(me, opp)
    if out is None:
        out = torch.empty((n, 2, 8, 8), dtype=torch.bfloat16, device=me.device)
    L = _lib.load()
    _lib.check(L.bz_reversi_planes(_lib.dptr(me), _lib.dptr(opp), _lib.dptr(out), n, _lib.stream_ptr()),
               "bz_reversi_planes")
    return out


# ------------------------------------------------------------------------------- tic-tac-toe
def ttt_legal_mask(x: torch.Tensor, o: torch.Tensor) -> torch.Tensor:
    mask = torch.empty_like(x)
    L = _lib.load()
    _lib.check(L.bz_ttt_legal_mask(_lib.dptr(x), _lib.dptr(o), _lib.dptr(mask), x.numel(), _lib.stream_ptr()),
               "bz_ttt_legal_mask")
    return mask


def ttt_apply(x: torch.Tensor, o: torch.Tensor, action: torch.Tensor, player: torch.Tensor):
    xo, oo = torch.empty_like(x), torch.empty_like(o)
    err = torch.empty(x.numel(), dtype=torch.uint8, device=x.device)
    L = _lib.load()
    _lib.check(L.bz_ttt_apply(_lib.dptr(x), _lib.dptr(o), _lib.dptr(action), _lib.dptr(player), _lib.dptr(xo),
                              _lib.dptr(oo), _lib.dptr(err), x.numel(), _lib.stream_ptr()), "bz_ttt_apply")
    return xo, oo, err


def ttt_terminal(x: torch.Tensor, o: torch.Tensor):
    over = torch.empty(x.numel(), dtype=torch.uint8, device=x.device)
    win = torch.empty(x.numel(), dtype=torch.int8, device=x.device)
    L = _lib.load()
    _lib.check(L.bz_ttt_terminal(_lib.dptr(x), _lib.dptr(o), _lib.dptr(over), _lib.dptr(win), x.numel(),
                                 _lib.stream_ptr()), "bz_ttt_terminal")
    return over, win


# ------------------------------------------------------------------------------- misc
def int32_microbench(blocks: int = 148 * 8, threads: int = 256, iters: int = 4096, variant: int = 0):
    """Runs the LOP3/SHF issue-rate microbenchmark once; returns integer instructions executed
    (per-thread count x threads).  Time it with CUDA events around the call."""
    import ctypes as C

    sink = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops = C.c_int64(0)
    L = _lib.load()
    _lib.check(L.bz_int32_microbench(_lib.dptr(sink), blocks, threads, iters, variant, C.byref(ops), _lib.stream_ptr()),
               "bz_int32_microbench")
    return ops.value * blocks * threads
