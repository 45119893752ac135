"""TEST INFRASTRUCTURE ONLY -- the *definition* of the sequential AlphaZero MCTS.

The reference (whaiproject/BetaZero) ships no MCTS, so **MCTS parity is
unpinned by the reference**.  This file fixes the semantics once, in pure Python, calling ONLY
the reference's board API for game logic:

    ReversiBoard.generate_possible_moves / make_move / is_game_over / get_score
        (src/reversi/game_logic/reversi_board.py:43-88)
    pass = board unchanged, player flips (src/reversi/game_logic/reversi_terminal.py:31-35)
    TicTacToeBoard.generate_possible_moves / make_move / is_game_over
        (src/tic_tac_toe/tic_tac_toe_board.py:23-43)
    canonical form = symbol * board (src/tic_tac_toe/players.py:85)

so it runs unchanged on the live reference classes (oracle/ref_shim.py) and on the C
restatement's wrappers (oracle/pyoracle.py).  oracle.c part 2 and the CUDA kernels
(betazero_b200/csrc/mcts.cu) both restate THIS file.

Frozen semantics
----------------
* Action ids are the engine's: Reversi ``a = row*8 + col`` for every board size, ``64`` = pass
  (legal iff the mover has no move and the game is not over); tic-tac-toe ``a = row*3 + col``.
  Edges of a node are ordered by ascending action id == the row-major order of
  ``generate_possible_moves`` (reversi_board.py:88), which is the tie-break order.
* One search = ``n_sims`` iterations of select -> (evaluate) -> expand+backup on an initially
  EMPTY tree; the first iteration expands the root (path length 0, nothing to back up).
* Edge statistics ``N:int32, W:float32, P:float32``.  ``Q = W / N`` if ``N > 0`` else ``0``.
* ``score = Q + ((c_puct * P) * sqrt(float32(n_node))) / float32(1 + N)`` -- every operation
  is a separately rounded float32 operation in exactly this order (no FMA);
  ``n_node = 1 + sum(child N)``.  argmax with strict ``>`` scanning ascending action ids.
* Expansion: evaluator ``f(me, opp) -> (w[A] >= 0, v)`` on the canonical (mover-relative)
  board; ``P_a = w_a / s`` with ``s`` the float32 sum of ``w`` over legal actions accumulated
  in ascending action order (``s == 0`` => uniform ``1/len(legal)``).  A no-move, not-over position is a normal node with the single
  edge ``pass`` (it is evaluated like any other node).
* Terminal leaf: ``v = winner * mover`` (get_score winner, reversi_board.py:68-76; TTT
  is_game_over winner); no evaluator call; nothing is expanded.
* Backup: ``v`` is the value for the side to move at the leaf; walking up, ``v = -v`` before
  each edge update (the sign flips every ply, pass plies included); ``N += 1; W = W + v``.
* Dirichlet noise off, one leaf per tree per iteration (no virtual loss).
* Tree reuse across moves is OPT-IN (``MCTS.advance``; every golden uses a fresh tree per move): after a move the
  search may continue on the subtree of the child the move leads to.  That subtree is kept iff the child has been
  expanded and the subtree's size, counted in the engine's arena units (``node_units``), is at most ``cap_units``;
  otherwise the next search starts on an empty tree.  A kept root keeps its statistics, and the next search ADDS its
  simulations to them (the root's visit total under the square root continues from the visits of the edge that led
  to it).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
M64 = (1 << 64) - 1


# --------------------------------------------------------------------------- evaluator
def _mix64(z: int) -> int:
    z &= M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def hash_eval(me: int, opp: int, salt: int, n_actions: int):
    """Parity-mode evaluator ("hash pseudo-net"): integer weights 1..32 and a value k/8, all
    exact in float32.  Same definition as orc_hash_eval (oracle.c) and bz_hash_eval (CUDA)."""
    h = _mix64((me * 0x9E3779B97F4A7C15 + _mix64(opp + 0xD1B54A32D192ED03) + salt) & M64)
    w = np.empty(n_actions, dtype=np.float32)
    for a in range(n_actions):
        w[a] = 1 + (_mix64((h + (a + 1) * 0x9E3779B97F4A7C15) & M64) >> 59)
    v = F32((((h >> 11) & 15) - 8) * 0.125)
    return w, v


# --------------------------------------------------------------------------- game adapters
class ReversiGame:
    """Adapter over a reference-compatible ReversiBoard class."""

    n_actions = 65
    PASS = 64

    def __init__(self, board_cls, size: int = 8):
        self.board_cls = board_cls
        self.size = size

    def initial(self):
        return self.board_cls(size=self.size), 1  # X (+1) moves first: reversi_terminal.py:14

    def terminal_value(self, board, player):
        if board.is_game_over():
            winner, _ = board.get_score()
            return F32(winner * player)
        return None

    def legal_actions(self, board, player):
        """Only called on non-terminal positions."""
        moves = board.generate_possible_moves(player)
        return [r * 8 + c for (r, c) in moves] if moves else [self.PASS]

    def next(self, board, player, action):
        if action == self.PASS:
            return board, -player
        return board.make_move(action >> 3, action & 7, player), -player

    def wire(self, board, player):
        g = np.asarray(board.board)
        me = opp = 0
        for r in range(self.size):
            for c in range(self.size):
                v = int(g[r, c])
                if v == player:
                    me |= 1 << (r * 8 + c)
                elif v == -player:
                    opp |= 1 << (r * 8 + c)
        return me, opp


class TicTacToeGame:
    """Adapter over a reference-compatible TicTacToeBoard class."""

    n_actions = 9

    def __init__(self, board_cls):
        self.board_cls = board_cls

    def initial(self):
        return self.board_cls(), 1  # tic_tac_toe.py:10

    def terminal_value(self, board, player):
        over, winner = board.is_game_over()
        return F32(winner * player) if over else None

    def legal_actions(self, board, player):
        return [r * 3 + c for (r, c) in board.generate_possible_moves()]

    def next(self, board, player, action):
        return board.make_move(action // 3, action % 3, player), -player

    def wire(self, board, player):
        g = np.asarray(board.board).reshape(-1)
        me = opp = 0
        for k in range(9):
            v = int(g[k])
            if v == player:
                me |= 1 << k
            elif v == -player:
                opp |= 1 << k
        return me, opp


# --------------------------------------------------------------------------- the search
class _Node:
    __slots__ = ("board", "player", "value", "actions", "P", "N", "W", "children", "visits")

    def __init__(self, board, player, value):
        self.board, self.player, self.value = board, player, value  # value != None => terminal
        self.actions = None  # None until expanded
        self.visits = 0  # descents that have entered the node (virtual-loss mode)


class MCTS:
    """Sequential single-tree PUCT search.  ``evaluator(me, opp) -> (w[A], v)``."""

    def __init__(self, game, c_puct: float = 1.25, evaluator=None):
        self.game = game
        self.c_puct = F32(c_puct)
        self.evaluator = evaluator
        self.root = None
        self._path = []
        self._leaf = None
        self.n_sims = 0
        self.sum_depth = 0

    def reset(self, board, player):
        self.root = _Node(board, player, self.game.terminal_value(board, player))
        self._leaf = None

    # -- select ------------------------------------------------------------------------
    def select(self):
        """Returns (status, me, opp): status 0 = needs evaluation, 1 = terminal."""
        g = self.game
        node, path = self.root, []
        while node.value is None and node.actions is not None:
            n_node = 1 + int(np.sum(node.N))
            sq = np.sqrt(F32(n_node))  # float32 sqrt, correctly rounded
            best, best_score = 0, None
            for i in range(len(node.actions)):
                n = int(node.N[i])
                q = node.W[i] / F32(n) if n > 0 else F32(0)
                u = self.c_puct * node.P[i]
                u = u * sq
                u = u / F32(1 + n)
                score = q + u
                if best_score is None or score > best_score:
                    best, best_score = i, score
            path.append((node, best))
            child = node.children[best]
            if child is None:
                nb, npl = g.next(node.board, node.player, node.actions[best])
                child = _Node(nb, npl, g.terminal_value(nb, npl))
                node.children[best] = child
                node = child
                break
            node = child
        self._path, self._leaf = path, node
        me, opp = g.wire(node.board, node.player)
        return (1 if node.value is not None else 0), me, opp

    # -- expand + backup ---------------------------------------------------------------
    def expand_backup(self, w, v):
        leaf = self._leaf
        if leaf.value is not None:
            value = F32(leaf.value)
        else:
            acts = self.game.legal_actions(leaf.board, leaf.player)
            s = F32(0)
            for a in acts:
                s = s + F32(w[a])
            leaf.actions = acts
            if s == 0:  # all-zero weights over the legal actions: uniform prior
                leaf.P = np.array([F32(1) / F32(len(acts))] * len(acts), dtype=np.float32)
            else:
                leaf.P = np.array([F32(w[a]) / s for a in acts], dtype=np.float32)
            leaf.N = np.zeros(len(acts), dtype=np.int32)
            leaf.W = np.zeros(len(acts), dtype=np.float32)
            leaf.children = [None] * len(acts)
            value = F32(v)
        for node, i in reversed(self._path):
            value = -value
            node.N[i] += 1
            node.W[i] = node.W[i] + value
        self.sum_depth += len(self._path)
        self.n_sims += 1
        self._leaf = None

    # -- tree reuse across moves (opt-in) -------------------------------------------------------
    def advance(self, action: int, board, player, cap_units: int | None = None) -> bool:
        """The game has played ``action`` from the root position and stands at ``(board, player)``.  Keep the subtree
        of the child that action leads to -- iff that child has been expanded and the subtree occupies at most
        ``cap_units`` arena units (None: no limit) -- as the tree of the next search; otherwise start an empty tree
        at the new position.  Returns True if the subtree was kept.

        The kept root's ``visits`` (virtual-loss mode: descents that have entered it) equal the N of the edge that led
        to it, and ``1 + sum(child N)`` of the one-leaf mode is the same number: the next search continues from it."""
        r = self.root
        child = None
        if r is not None and r.actions is not None and action in r.actions:
            child = r.children[r.actions.index(action)]
        if child is not None and child.value is None and child.actions is not None and (
                cap_units is None or subtree_units(child) <= cap_units):
            self.root = child
            self._leaf = None
            return True
        self.reset(board, player)
        return False

    def run(self, n_sims: int):
        for _ in range(n_sims):
            status, me, opp = self.select()
            if status == 0:
                w, v = self.evaluator(me, opp)
            else:
                w, v = None, F32(0)
            self.expand_backup(w, v)

    # -- K leaves per iteration with virtual loss -------------------------------------------
    # An iteration makes K descents one after the other (slots 0..K-1), evaluates the K leaves, then
    # expands / backs them up in slot order.  A descent leaves a virtual loss on every edge it walks
    # (N += 1, W = W - 1 in float32) so the following descents of the iteration are steered elsewhere.
    # The visit count under the square root is the number of descents that have ENTERED the node
    # before this one; for K = 1 that is exactly 1 + sum(child N) of select() above.  Two descents may
    # end on the same unexpanded leaf: the first backup expands it, later ones only back up their value.
    # Backup turns the virtual visit into a real one: W = (W + 1) + value (two float32 adds), N unchanged.
    def select_vl(self, slot: int):
        g = self.game
        if not hasattr(self, "_vl"):
            self._vl = {}
        node, path = self.root, []
        while True:
            n_node = node.visits
            node.visits += 1
            if node.value is not None or node.actions is None:
                break
            sq = np.sqrt(F32(n_node))
            best, best_score = 0, None
            for i in range(len(node.actions)):
                n = int(node.N[i])
                q = node.W[i] / F32(n) if n > 0 else F32(0)
                u = self.c_puct * node.P[i]
                u = u * sq
                u = u / F32(1 + n)
                score = q + u
                if best_score is None or score > best_score:
                    best, best_score = i, score
            path.append((node, best))
            node.N[best] += 1
            node.W[best] = node.W[best] - F32(1)
            child = node.children[best]
            if child is None:
                nb, npl = g.next(node.board, node.player, node.actions[best])
                child = _Node(nb, npl, g.terminal_value(nb, npl))
                node.children[best] = child
            node = child
        self._vl[slot] = (path, node)
        me, opp = g.wire(node.board, node.player)
        return (1 if node.value is not None else 0), me, opp

    def expand_backup_vl(self, slot: int, w, v):
        path, leaf = self._vl.pop(slot)
        if leaf.value is not None:
            value = F32(leaf.value)
        elif leaf.actions is not None:
            value = F32(v)  # expanded by an earlier slot of this iteration
        else:
            acts = self.game.legal_actions(leaf.board, leaf.player)
            s = F32(0)
            for a in acts:
                s = s + F32(w[a])
            leaf.actions = acts
            if s == 0:
                leaf.P = np.array([F32(1) / F32(len(acts))] * len(acts), dtype=np.float32)
            else:
                leaf.P = np.array([F32(w[a]) / s for a in acts], dtype=np.float32)
            leaf.N = np.zeros(len(acts), dtype=np.int32)
            leaf.W = np.zeros(len(acts), dtype=np.float32)
            leaf.children = [None] * len(acts)
            value = F32(v)
        for node, i in reversed(path):
            value = -value
            node.W[i] = (node.W[i] + F32(1)) + value
        self.sum_depth += len(path)
        self.n_sims += 1

    def run_vl(self, n_sims: int, leaves: int):
        assert n_sims % leaves == 0
        for _ in range(n_sims // leaves):
            evs = []
            for j in range(leaves):
                status, me, opp = self.select_vl(j)
                evs.append(self.evaluator(me, opp) if status == 0 else (None, F32(0)))
            for j in range(leaves):
                self.expand_backup_vl(j, *evs[j])

    # -- results -----------------------------------------------------------------------
    def root_stats(self):
        """(visit counts int32[A], W float32[A], P float32[A]) scattered by action."""
        A = self.game.n_actions
        cnt = np.zeros(A, dtype=np.int32)
        W = np.zeros(A, dtype=np.float32)
        P = np.zeros(A, dtype=np.float32)
        r = self.root
        if r is not None and r.actions is not None:
            for i, a in enumerate(r.actions):
                cnt[a], W[a], P[a] = r.N[i], r.W[i], r.P[i]
        return cnt, W, P


def node_units(n_edges: int) -> int:
    """Arena units (32 bytes) of an expanded node with ``n_edges`` edges in the engine's tree layout: an 8-word header
    and four words per edge, rounded up to whole units (include/betazero_b200.h)."""
    return (8 + 4 * n_edges + 7) >> 3


def subtree_units(node) -> int:
    """Arena units of the expanded nodes of the subtree below (and including) ``node``; unexpanded and terminal
    children occupy nothing."""
    total, stack = 0, [node]
    while stack:
        nd = stack.pop()
        if nd is None or nd.actions is None:
            continue
        total += node_units(len(nd.actions))
        stack.extend(nd.children)
    return total


def policy_from_counts(counts):
    """pi = N / sum(N) in float32 (zeros if nothing was visited)."""
    counts = np.asarray(counts, dtype=np.int32)
    tot = int(counts.sum())
    if tot == 0:
        return np.zeros(counts.shape, dtype=np.float32)
    return (counts.astype(np.float32) / F32(tot)).astype(np.float32)


def pick_move(counts) -> int:
    """Deterministic move choice: most visited action, lowest action id on ties."""
    return int(np.argmax(np.asarray(counts)))  # np.argmax returns the first maximum


def self_play_game(game, n_sims: int, c_puct: float, evaluator, max_plies: int = 200, leaves: int = 1, reuse: bool = False,
                   cap_units: int | None = None):
    """One deterministic self-play game in the order of the reference episode loops
    (reversi_terminal.py:16-38 / tic_tac_toe.py:13-34): search, move (or pass), terminal test,
    flip player.  Returns a list of (me, opp, player, counts, action) per ply and the winner.
    ``leaves`` > 1: every search makes that many virtual-loss descents per iteration (MCTS.run_vl).
    ``reuse``: the tree of a move continues on the subtree of the move played (MCTS.advance) instead of starting empty."""
    board, player = game.initial()
    history = []
    m = None
    while game.terminal_value(board, player) is None and len(history) < max_plies:
        if m is None or not reuse:
            m = MCTS(game, c_puct, evaluator)
            m.reset(board, player)
        if leaves > 1:
            m.run_vl(n_sims, leaves)
        else:
            m.run(n_sims)
        cnt, _, _ = m.root_stats()
        a = pick_move(cnt)
        me, opp = game.wire(board, player)
        history.append((me, opp, player, cnt, a))
        board, player = game.next(board, player, a)
        if reuse:
            m.advance(a, board, player, cap_units)
    tv = game.terminal_value(board, player)
    winner = 0 if tv is None else int(tv) * player
    return history, winner
