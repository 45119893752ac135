"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/liboracle.so (oracle.c).

Exposes
* ``OracleReversiBoard`` / ``OracleTicTacToeBoard``: the reference board API
  (reversi_board.py:3-88, tic_tac_toe_board.py:3-43) on top of the C restatement, so the
  Python MCTS definition (mcts_ref.py) and reference-style episode loops run on the GPU box,
  where /root/reference does not exist;
* batched wire-format helpers (``legal_mask``, ``apply``, ``terminal`` ...) used as the
  checker of the CUDA env kernels;
* ``OracleTree``: the C sequential MCTS (oracle.c part 2), ``search_hash``: whole searches with
  the hash evaluator, multi-threaded over trees (CPU baseline).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def _src_hash() -> str:
    import hashlib

    h = hashlib.sha256()
    for name in ("oracle.c", "Makefile"):
        with open(os.path.join(_HERE, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so (gcc, see Makefile).  Staleness is decided by a content hash kept beside the
    library (file times do not survive a copy to another box) and the build runs under an inter-process lock, so the
    ranks of a torchrun job never compile -- or load a half-written library -- at the same time."""
    import fcntl

    stamp = _SO + ".srchash"

    def stale():
        try:
            return not os.path.exists(_SO) or open(stamp).read().strip() != _src_hash()
        except OSError:
            return True

    if not force and not stale():
        return _SO
    with open(_SO + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or stale():
                subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
                with open(stamp + ".tmp", "w") as f:
                    f.write(_src_hash())
                os.replace(stamp + ".tmp", stamp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


def lib():
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not os.path.exists(_SO):
                raise
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


_p8 = C.POINTER(C.c_int8)
_pu8 = C.POINTER(C.c_uint8)
_pu16 = C.POINTER(C.c_uint16)
_pu64 = C.POINTER(C.c_uint64)
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)
_pf = C.POINTER(C.c_float)
_pint = C.POINTER(C.c_int)


def _declare(L):
    L.orc_reversi_init.argtypes = [_p8, C.c_int]
    L.orc_reversi_is_valid_move.argtypes = [_p8, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_reversi_make_move.argtypes = [_p8, C.c_int, C.c_int, C.c_int, C.c_int, _p8]
    L.orc_reversi_generate_possible_moves.argtypes = [_p8, C.c_int, C.c_int, _pint]
    L.orc_reversi_is_game_over.argtypes = [_p8, C.c_int]
    L.orc_reversi_get_score.argtypes = [_p8, C.c_int, _pint, _pint]
    L.orc_ttt_is_valid_move.argtypes = [_p8, C.c_int, C.c_int]
    L.orc_ttt_make_move.argtypes = [_p8, C.c_int, C.c_int, C.c_int, _p8]
    L.orc_ttt_is_game_over.argtypes = [_p8, _pint]
    L.orc_ttt_generate_possible_moves.argtypes = [_p8, _pint]
    L.orc_reversi_batch_legal_mask.argtypes = [_pu64, _pu64, _pu64, C.c_int64, C.c_int]
    L.orc_reversi_batch_apply.argtypes = [_pu64, _pu64, _pu8, _pu64, _pu64, _pu8, C.c_int64, C.c_int]
    L.orc_reversi_batch_terminal.argtypes = [_pu64, _pu64, _pu8, _p8, _pu8, _pu8, C.c_int64, C.c_int]
    L.orc_ttt_batch_legal_mask.argtypes = [_pu16, _pu16, _pu16, C.c_int64]
    L.orc_ttt_batch_apply.argtypes = [_pu16, _pu16, _pu8, _p8, _pu16, _pu16, _pu8, C.c_int64]
    L.orc_ttt_batch_terminal.argtypes = [_pu16, _pu16, _pu8, _p8, C.c_int64]
    L.orc_reversi_random_playout.argtypes = [C.c_int, C.c_uint64, C.c_int, _p8, _pint, _pint]
    L.orc_mcts_new.argtypes = [C.c_int, C.c_int, C.c_float]
    L.orc_mcts_new.restype = C.c_void_p
    L.orc_mcts_free.argtypes = [C.c_void_p]
    L.orc_mcts_reset.argtypes = [C.c_void_p, _p8, C.c_int]
    L.orc_mcts_select.argtypes = [C.c_void_p, _pu64, _pu64, _pint]
    L.orc_mcts_expand_backup.argtypes = [C.c_void_p, _pf, C.c_float]
    L.orc_mcts_root_stats.argtypes = [C.c_void_p, _pi32, _pf, _pf]
    L.orc_mcts_counters.argtypes = [C.c_void_p, _pi64]
    L.orc_hash_eval.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, _pf, _pf]
    L.orc_mcts_search_hash.argtypes = [C.c_int, C.c_int, C.c_float, C.c_uint64, _pu64, _pu64, C.c_int64, C.c_int,
                                       _pi32, _pf, _pf, _pi64]
    L.orc_num_threads.restype = C.c_int
    _pp = C.POINTER(C.c_void_p)
    L.orc_mcts_batch_reset_wire.argtypes = [_pp, C.c_int64, _pu64, _pu64]
    L.orc_mcts_batch_select.argtypes = [_pp, C.c_int64, _pu64, _pu64, _pu8, _pf]
    L.orc_mcts_batch_expand_backup.argtypes = [_pp, C.c_int64, _pf, _pf, C.c_int]
    L.orc_mcts_batch_root_counts.argtypes = [_pp, C.c_int64, _pi32, C.c_int]
    L.orc_mcts_select_vl.argtypes = [C.c_void_p, C.c_int, _pu64, _pu64, _pint]
    L.orc_mcts_expand_backup_vl.argtypes = [C.c_void_p, C.c_int, _pf, C.c_float]
    L.orc_mcts_search_hash_vl.argtypes = [C.c_int, C.c_int, C.c_float, C.c_uint64, _pu64, _pu64, C.c_int64, C.c_int, C.c_int,
                                          _pi32, _pf, _pf, _pi64]
    L.orc_mcts_advance.argtypes = [C.c_void_p, C.c_int, _p8, C.c_int, C.c_int64]
    L.orc_mcts_advance.restype = C.c_int
    L.orc_mcts_play_hash.argtypes = [C.c_int, C.c_int, C.c_float, C.c_uint64, _pu64, _pu64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int64, _pi32, _pu8, _pu8]
    L.orc_mcts_batch_select_vl.argtypes = [_pp, C.c_int64, C.c_int, _pu64, _pu64, _pu8, _pf]
    L.orc_mcts_batch_expand_backup_vl.argtypes = [_pp, C.c_int64, C.c_int, _pf, _pf, C.c_int]


def _ptr(a: np.ndarray, ty):
    return a.ctypes.data_as(ty)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# --------------------------------------------------------------------------- board classes
class OracleReversiBoard:
    """Reference ReversiBoard API (reversi_board.py:3-88) on the C ray-walk restatement."""

    def __init__(self, board=None, size=8):
        if board is None:
            self.size = size
            g = np.zeros((size, size), dtype=np.int8)
            lib().orc_reversi_init(_ptr(g, _p8), size)
            self.board = g.astype(np.int64)
        else:  # copy-ctor takes another board object (reversi_board.py:13-14)
            self.board = np.copy(board.board)
            self.size = int(board.size)

    def _g(self):
        return _c(self.board, np.int8)

    def is_valid_move(self, row, col, player):
        return bool(lib().orc_reversi_is_valid_move(_ptr(self._g(), _p8), self.size, int(row), int(col), int(player)))

    def make_move(self, row, col, player):
        out = np.empty((self.size, self.size), dtype=np.int8)
        rc = lib().orc_reversi_make_move(_ptr(self._g(), _p8), self.size, int(row), int(col), int(player),
                                         _ptr(out, _p8))
        if rc != 0:
            raise ValueError("Invalid move")
        nb = OracleReversiBoard.__new__(OracleReversiBoard)
        nb.size, nb.board = self.size, out.astype(np.int64)
        return nb

    def is_game_over(self):
        return bool(lib().orc_reversi_is_game_over(_ptr(self._g(), _p8), self.size))

    def get_score(self, print_result=False):
        c1, c2 = C.c_int(), C.c_int()
        w = lib().orc_reversi_get_score(_ptr(self._g(), _p8), self.size, C.byref(c1), C.byref(c2))
        return w, (c1.value, c2.value)

    def generate_possible_moves(self, player):
        mv = (C.c_int * 64)()
        n = lib().orc_reversi_generate_possible_moves(_ptr(self._g(), _p8), self.size, int(player), mv)
        return [(mv[i] // self.size, mv[i] % self.size) for i in range(n)]


class OracleTicTacToeBoard:
    """Reference TicTacToeBoard API (tic_tac_toe_board.py:3-43) on the C restatement."""

    def __init__(self, board=None):
        self.board = np.zeros((3, 3), dtype=np.int64) if board is None else np.copy(board)

    def _g(self):
        return _c(self.board, np.int8)

    def is_valid_move(self, row, col):
        return bool(lib().orc_ttt_is_valid_move(_ptr(self._g(), _p8), int(row), int(col)))

    def make_move(self, row, col, player):
        out = np.empty((3, 3), dtype=np.int8)
        if lib().orc_ttt_make_move(_ptr(self._g(), _p8), int(row), int(col), int(player), _ptr(out, _p8)) != 0:
            raise ValueError("Invalid move")
        return OracleTicTacToeBoard(out.astype(np.int64))

    def is_game_over(self):
        w = C.c_int()
        over = lib().orc_ttt_is_game_over(_ptr(self._g(), _p8), C.byref(w))
        return (True, w.value) if over else (False, None)

    def generate_possible_moves(self):
        mv = (C.c_int * 9)()
        n = lib().orc_ttt_generate_possible_moves(_ptr(self._g(), _p8), mv)
        return [(mv[i] // 3, mv[i] % 3) for i in range(n)]


# --------------------------------------------------------------------------- wire-format helpers
def grid_to_wire(grid, player=1):
    """int grid (+1/-1/0) [size,size] -> (me, opp) python ints, bit = row*8+col."""
    g = np.asarray(grid)
    me = opp = 0
    for r in range(g.shape[0]):
        for c in range(g.shape[1]):
            v = int(g[r, c])
            if v == player:
                me |= 1 << (r * 8 + c)
            elif v == -player:
                opp |= 1 << (r * 8 + c)
    return me, opp


def wire_to_grid(me, opp, size=8, player=1):
    g = np.zeros((size, size), dtype=np.int64)
    for r in range(size):
        for c in range(size):
            b = r * 8 + c
            if (int(me) >> b) & 1:
                g[r, c] = player
            elif (int(opp) >> b) & 1:
                g[r, c] = -player
    return g


def legal_mask(me, opp, size=8):
    me, opp = _c(me, np.uint64), _c(opp, np.uint64)
    out = np.empty_like(me)
    lib().orc_reversi_batch_legal_mask(_ptr(me, _pu64), _ptr(opp, _pu64), _ptr(out, _pu64), me.size, size)
    return out


def apply(me, opp, action, size=8):
    me, opp, action = _c(me, np.uint64), _c(opp, np.uint64), _c(action, np.uint8)
    mo, oo, err = np.empty_like(me), np.empty_like(me), np.empty(me.size, dtype=np.uint8)
    lib().orc_reversi_batch_apply(_ptr(me, _pu64), _ptr(opp, _pu64), _ptr(action, _pu8), _ptr(mo, _pu64),
                                  _ptr(oo, _pu64), _ptr(err, _pu8), me.size, size)
    return mo, oo, err


def terminal(me, opp, size=8):
    me, opp = _c(me, np.uint64), _c(opp, np.uint64)
    n = me.size
    over, win = np.empty(n, np.uint8), np.empty(n, np.int8)
    cm, co = np.empty(n, np.uint8), np.empty(n, np.uint8)
    lib().orc_reversi_batch_terminal(_ptr(me, _pu64), _ptr(opp, _pu64), _ptr(over, _pu8), _ptr(win, _p8),
                                     _ptr(cm, _pu8), _ptr(co, _pu8), n, size)
    return over, win, cm, co


def ttt_legal_mask(x, o):
    x, o = _c(x, np.uint16), _c(o, np.uint16)
    out = np.empty_like(x)
    lib().orc_ttt_batch_legal_mask(_ptr(x, _pu16), _ptr(o, _pu16), _ptr(out, _pu16), x.size)
    return out


def ttt_apply(x, o, action, player):
    x, o = _c(x, np.uint16), _c(o, np.uint16)
    action, player = _c(action, np.uint8), _c(player, np.int8)
    xo, oo, err = np.empty_like(x), np.empty_like(x), np.empty(x.size, np.uint8)
    lib().orc_ttt_batch_apply(_ptr(x, _pu16), _ptr(o, _pu16), _ptr(action, _pu8), _ptr(player, _p8),
                              _ptr(xo, _pu16), _ptr(oo, _pu16), _ptr(err, _pu8), x.size)
    return xo, oo, err


def ttt_terminal(x, o):
    x, o = _c(x, np.uint16), _c(o, np.uint16)
    over, win = np.empty(x.size, np.uint8), np.empty(x.size, np.int8)
    lib().orc_ttt_batch_terminal(_ptr(x, _pu16), _ptr(o, _pu16), _ptr(over, _pu8), _ptr(win, _p8), x.size)
    return over, win


def random_playout(size, seed, max_plies=-1):
    """reference episode loop with random players; returns (grid int64, player_to_move, plies, passes)."""
    g = np.zeros((size, size), dtype=np.int8)
    pl, ps = C.c_int(), C.c_int()
    n = lib().orc_reversi_random_playout(size, seed, max_plies, _ptr(g, _p8), C.byref(pl), C.byref(ps))
    return g.astype(np.int64), pl.value, n, ps.value


# --------------------------------------------------------------------------- boards for tests / bench
def synthetic_boards(n, seed=0):
    """Config-2 "set A" (SURVEY.md section 8d): per board p_empty ~ U[0.05, 0.9], each cell iid
    {empty, me, opp} with (p_empty, (1-p_empty)/2, (1-p_empty)/2).  Returns (me, opp) uint64[n]."""
    rng = np.random.default_rng(seed)
    p_empty = rng.uniform(0.05, 0.9, size=(n, 1))
    u = rng.random((n, 64))
    v = rng.random((n, 64)) < 0.5
    occ = u >= p_empty
    bits = (np.uint64(1) << np.arange(64, dtype=np.uint64))[None, :]
    me = np.bitwise_or.reduce(np.where(occ & v, bits, np.uint64(0)), axis=1)
    opp = np.bitwise_or.reduce(np.where(occ & ~v, bits, np.uint64(0)), axis=1)
    return me.astype(np.uint64), opp.astype(np.uint64)


def playout_boards(n, seed=0, size=8):
    """Config-2 "set B": reachable boards from seeded random playouts of 0..60 loop iterations
    (reference loop, reversi_terminal.py:16-38).  Returns mover-relative (me, opp) uint64[n]."""
    rng = np.random.default_rng(seed)
    plies = rng.integers(0, size * size - 3, size=n)
    me = np.empty(n, np.uint64)
    opp = np.empty(n, np.uint64)
    for i in range(n):
        g, pl, _, _ = random_playout(size, int(seed * 1_000_003 + i), int(plies[i]))
        m, o = grid_to_wire(g, pl)
        me[i], opp[i] = m, o
    return me, opp


# --------------------------------------------------------------------------- MCTS
GAME_REVERSI, GAME_TTT = 0, 1


def hash_eval(me, opp, salt, n_actions):
    w = np.empty(n_actions, np.float32)
    v = C.c_float()
    lib().orc_hash_eval(int(me), int(opp), int(salt), n_actions, _ptr(w, _pf), C.byref(v))
    return w, np.float32(v.value)


class OracleTree:
    """C sequential MCTS, stepwise (select / expand_backup) like the GPU API."""

    def __init__(self, game=GAME_REVERSI, size=8, c_puct=1.25):
        self.game, self.size = game, (3 if game == GAME_TTT else size)
        self.n_actions = 9 if game == GAME_TTT else 65
        self._t = lib().orc_mcts_new(game, size, c_puct)

    def __del__(self):
        if getattr(self, "_t", None):
            lib().orc_mcts_free(self._t)
            self._t = None

    def reset(self, grid, player):
        g = _c(grid, np.int8)
        lib().orc_mcts_reset(self._t, _ptr(g, _p8), int(player))

    def reset_wire(self, me, opp):
        """root given mover-relative; the mover is encoded as +1."""
        if self.game == GAME_TTT:
            g = np.zeros(9, np.int8)
            for k in range(9):
                g[k] = 1 if (int(me) >> k) & 1 else (-1 if (int(opp) >> k) & 1 else 0)
        else:
            g = wire_to_grid(me, opp, self.size)
        self.reset(g, 1)

    def select(self):
        me, opp, d = C.c_uint64(), C.c_uint64(), C.c_int()
        st = lib().orc_mcts_select(self._t, C.byref(me), C.byref(opp), C.byref(d))
        return st, me.value, opp.value, d.value

    def expand_backup(self, w, v):
        w = _c(w if w is not None else np.zeros(self.n_actions), np.float32)
        lib().orc_mcts_expand_backup(self._t, _ptr(w, _pf), float(v))

    # virtual-loss mode: several descents in flight (slots), oracle.c Part 2d
    def select_vl(self, slot):
        me, opp, d = C.c_uint64(), C.c_uint64(), C.c_int()
        st = lib().orc_mcts_select_vl(self._t, int(slot), C.byref(me), C.byref(opp), C.byref(d))
        return st, me.value, opp.value, d.value

    def expand_backup_vl(self, slot, w, v):
        w = _c(w if w is not None else np.zeros(self.n_actions), np.float32)
        lib().orc_mcts_expand_backup_vl(self._t, int(slot), _ptr(w, _pf), float(v))

    def root_stats(self):
        cnt = np.zeros(self.n_actions, np.int32)
        W = np.zeros(self.n_actions, np.float32)
        P = np.zeros(self.n_actions, np.float32)
        lib().orc_mcts_root_stats(self._t, _ptr(cnt, _pi32), _ptr(W, _pf), _ptr(P, _pf))
        return cnt, W, P

    def counters(self):
        c = np.zeros(4, np.int64)
        lib().orc_mcts_counters(self._t, _ptr(c, _pi64))
        return dict(nodes=int(c[0]), edges=int(c[1]), sum_depth=int(c[2]), sims=int(c[3]))


def search_hash(me, opp, n_sims, game=GAME_REVERSI, size=8, c_puct=1.25, salt=0, leaves=1):
    """n_sims-iteration searches from every (me, opp) root with the hash evaluator, OpenMP over
    trees.  Returns (counts int32[n,A], W f32[n,A], P f32[n,A], counters dict).  ``leaves`` > 1: the
    virtual-loss mode, ``leaves`` descents per iteration (n_sims must be a multiple)."""
    me, opp = _c(me, np.uint64), _c(opp, np.uint64)
    A = 9 if game == GAME_TTT else 65
    n = me.size
    cnt = np.zeros((n, A), np.int32)
    W = np.zeros((n, A), np.float32)
    P = np.zeros((n, A), np.float32)
    c = np.zeros(4, np.int64)
    if leaves > 1:
        if n_sims % leaves:
            raise ValueError("n_sims must be a multiple of leaves")
        lib().orc_mcts_search_hash_vl(game, size, c_puct, salt, _ptr(me, _pu64), _ptr(opp, _pu64), n, n_sims, leaves,
                                      _ptr(cnt, _pi32), _ptr(W, _pf), _ptr(P, _pf), _ptr(c, _pi64))
    else:
        lib().orc_mcts_search_hash(game, size, c_puct, salt, _ptr(me, _pu64), _ptr(opp, _pu64), n, n_sims,
                                   _ptr(cnt, _pi32), _ptr(W, _pf), _ptr(P, _pf), _ptr(c, _pi64))
    return cnt, W, P, dict(nodes=int(c[0]), edges=int(c[1]), sum_depth=int(c[2]), sims=int(c[3]))


def play_hash(me, opp, n_sims, plies, game=GAME_REVERSI, size=8, c_puct=1.25, salt=0, leaves=1, reuse=False, cap_units=-1):
    """``plies`` moves of deterministic play from every (me, opp) root with the hash evaluator (orc_mcts_play_hash): a
    search of n_sims simulations per move, the most visited action is played, and -- ``reuse`` -- the next move's tree
    is the kept subtree of the move played (MCTS.advance, at most ``cap_units`` arena units; -1 = no limit).
    Returns (counts int32[n, plies, A], actions uint8[n, plies] (255 after the game's end), kept uint8[n, plies])."""
    me, opp = _c(me, np.uint64), _c(opp, np.uint64)
    A = 9 if game == GAME_TTT else 65
    n = me.size
    if n_sims % leaves:
        raise ValueError("n_sims must be a multiple of leaves")
    cnt = np.zeros((n, plies, A), np.int32)
    act = np.zeros((n, plies), np.uint8)
    kept = np.zeros((n, plies), np.uint8)
    lib().orc_mcts_play_hash(game, size, c_puct, salt, _ptr(me, _pu64), _ptr(opp, _pu64), n, n_sims, leaves, plies,
                             1 if reuse else 0, int(cap_units), _ptr(cnt, _pi32), _ptr(act, _pu8), _ptr(kept, _pu8))
    return cnt, act, kept


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every host core."""
    lib().orc_set_num_threads(int(n))


class OracleForest:
    """Many C trees stepped in lockstep (OpenMP over trees): the CPU baseline's counterpart of the
    GPU's BatchedMCTS, usable with any evaluator (e.g. the same PyTorch net on the CPU)."""

    def __init__(self, n, game=GAME_REVERSI, size=8, c_puct=1.25, leaves=1):
        self.n, self.game, self.leaves = int(n), game, int(leaves)
        self.n_actions = 9 if game == GAME_TTT else 65
        self._trees = [lib().orc_mcts_new(game, size, c_puct) for _ in range(self.n)]
        self._arr = (C.c_void_p * self.n)(*self._trees)
        rows = self.n * self.leaves  # leaf arrays are slot-major [leaves][n]
        self.leaf_me = np.zeros(rows, np.uint64)
        self.leaf_opp = np.zeros(rows, np.uint64)
        self.status = np.zeros(rows, np.uint8)
        self.planes = np.zeros((rows, 2, 8, 8), np.float32)

    def __del__(self):
        for t in getattr(self, "_trees", []):
            lib().orc_mcts_free(t)
        self._trees = []

    def reset(self, me, opp):
        me, opp = _c(me, np.uint64), _c(opp, np.uint64)
        lib().orc_mcts_batch_reset_wire(self._arr, self.n, _ptr(me, _pu64), _ptr(opp, _pu64))

    def select(self):
        if self.leaves > 1:
            lib().orc_mcts_batch_select_vl(self._arr, self.n, self.leaves, _ptr(self.leaf_me, _pu64),
                                           _ptr(self.leaf_opp, _pu64), _ptr(self.status, _pu8), _ptr(self.planes, _pf))
        else:
            lib().orc_mcts_batch_select(self._arr, self.n, _ptr(self.leaf_me, _pu64), _ptr(self.leaf_opp, _pu64),
                                        _ptr(self.status, _pu8), _ptr(self.planes, _pf))

    def expand_backup(self, w, v):
        w, v = _c(w, np.float32), _c(v, np.float32)
        if self.leaves > 1:
            lib().orc_mcts_batch_expand_backup_vl(self._arr, self.n, self.leaves, _ptr(w, _pf), _ptr(v, _pf), self.n_actions)
        else:
            lib().orc_mcts_batch_expand_backup(self._arr, self.n, _ptr(w, _pf), _ptr(v, _pf), self.n_actions)

    def root_counts(self):
        cnt = np.zeros((self.n, self.n_actions), np.int32)
        lib().orc_mcts_batch_root_counts(self._arr, self.n, _ptr(cnt, _pi32), self.n_actions)
        return cnt
