"""Batched MCTS host side: tree pools in HBM + the lockstep search loop (kernels K5-K8).

The reference has no MCTS; the search serves its
``Player.get_move(board) -> (row, col)`` contract (reversi_players.py:5-8, players.py:6-9) and
follows the semantics frozen in oracle/mcts_ref.py.  One search of ``n_sims`` iterations is

    reset -> select -> [evaluate -> step]*(n_sims-1) -> evaluate -> expand_backup

where ``step`` = expand+backup of iteration i fused with the select+gather of iteration i+1 in
ONE kernel launch, and ``[evaluate -> step]`` blocks are replayed from a CUDA graph.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import BzTreePools

GAME_REVERSI, GAME_TTT = 0, 1
LEAF_EVAL, LEAF_TERMINAL, LEAF_ERROR = 0, 1, 2
MAX_EDGE_CAP = 0x7FFF0
_MAX_BRANCH = {GAME_REVERSI: 33, GAME_TTT: 9}  # 33 = most legal moves of any 8x8 Reversi position


class TreePools:
    """Caller-owned SoA pools for ``n_trees`` concurrent trees (``bz_tree_pools``).

    ``sims_cap``: the largest number of iterations a search will run between resets; every
    iteration expands at most one node, and a node has at most 33 (Reversi) / 9 (TTT) edges, so
    ``edge_cap = sims_cap * max_branch`` can never overflow.  A smaller ``edge_cap`` may be passed
    (mean branching is ~8-10); overflow is detected and raised, never silent.
    """

    def __init__(self, n_trees: int, sims_cap: int, game: int = GAME_REVERSI, board_size: int = 8,
                 c_puct: float = 1.25, edge_cap: int | None = None, max_depth: int | None = None, device="cuda"):
        if game not in (GAME_REVERSI, GAME_TTT):
            raise ValueError("game must be GAME_REVERSI or GAME_TTT")
        self.game, self.board_size = game, (3 if game == GAME_TTT else board_size)
        self.n_trees, self.sims_cap = int(n_trees), int(sims_cap)
        self.n_actions = 9 if game == GAME_TTT else 65
        self.c_puct = float(c_puct)
        if edge_cap is None:
            edge_cap = min(self.sims_cap * _MAX_BRANCH[game], MAX_EDGE_CAP)
        if edge_cap > MAX_EDGE_CAP:
            raise ValueError(f"edge_cap {edge_cap} exceeds {MAX_EDGE_CAP}")
        self.edge_cap = int(edge_cap)
        self.max_depth = int(max_depth or (16 if game == GAME_TTT else 128))
        self.device = torch.device(device)
        B, E, D = self.n_trees, self.n_trees * self.edge_cap, self.n_trees * self.max_depth
        dev = self.device

        def e(n, dt):
            return torch.empty(max(n, 1), dtype=dt, device=dev)

        self.root_me, self.root_opp = e(B, torch.int64), e(B, torch.int64)
        # valid "empty tree" state from the start, so no kernel ever walks uninitialised memory
        self.root_meta = torch.full((max(B, 1),), -8192, dtype=torch.int32, device=dev)  # meta(UNEXPANDED)
        self.edge_count, self.sim_count = torch.zeros(max(B, 1), dtype=torch.int32, device=dev), e(B, torch.int32)
        self.depth_sum, self.error = e(B, torch.int32), torch.zeros(max(B, 1), dtype=torch.int32, device=dev)
        self.edge_N, self.edge_W, self.edge_P = e(E, torch.int32), e(E, torch.float32), e(E, torch.float32)
        self.edge_meta = e(E, torch.int32)
        self.edge_me, self.edge_opp = e(E, torch.int64), e(E, torch.int64)
        self.path, self.path_len = e(D, torch.int32), e(B, torch.int32)
        self.leaf_me, self.leaf_opp, self.leaf_mask = e(B, torch.int64), e(B, torch.int64), e(B, torch.int64)
        self.leaf_status = torch.full((max(B, 1),), LEAF_ERROR, dtype=torch.uint8, device=dev)
        self.leaf_value = e(B, torch.float32)
        if game == GAME_TTT:
            self.leaf_planes = torch.zeros((max(B, 1), 9), dtype=torch.bfloat16, device=dev)
        else:
            self.leaf_planes = torch.zeros((max(B, 1), 2, 8, 8), dtype=torch.bfloat16, device=dev)
        s = BzTreePools()
        s.game, s.board_size, s.n_trees, s.n_actions = game, self.board_size if game == GAME_REVERSI else 8, B, self.n_actions
        s.edge_cap, s.max_depth, s.c_puct, s.reserved = self.edge_cap, self.max_depth, self.c_puct, 0
        for name, _ in BzTreePools._fields_[8:]:
            setattr(s, name, getattr(self, name).data_ptr())
        self.c_struct = s
        self._ref = C.byref(s)

    # bytes of HBM held by the pools
    def nbytes(self) -> int:
        return sum(getattr(self, n).numel() * getattr(self, n).element_size() for n, _ in BzTreePools._fields_[8:])


class HashEvaluator:
    """Parity-mode evaluator: the integer-hash pseudo-net (``bz_hash_eval``), exact in fp32 and
    identical to oracle/mcts_ref.py:hash_eval."""

    def __init__(self, salt: int = 0):
        self.salt = int(salt)

    def __call__(self, pools: TreePools, out_w: torch.Tensor, out_v: torch.Tensor) -> None:
        L = _lib.load()
        _lib.check(L.bz_hash_eval(_lib.dptr(pools.leaf_me), _lib.dptr(pools.leaf_opp), self.salt, pools.n_actions,
                                  _lib.dptr(out_w), _lib.dptr(out_v), pools.n_trees, _lib.stream_ptr()),
                   "bz_hash_eval")


class NetEvaluator:
    """PyTorch policy/value net on the gathered bf16 leaf planes.  The net returns
    ``(logits [B, A], value [B])``; prior weights are ``softmax(logits)`` in fp32 (the tree kernel
    renormalises them over the legal actions)."""

    def __init__(self, net: torch.nn.Module):
        self.net = net

    @torch.no_grad()
    def __call__(self, pools: TreePools, out_w: torch.Tensor, out_v: torch.Tensor) -> None:
        logits, value = self.net(pools.leaf_planes)
        torch.softmax(logits.float(), dim=-1, out=out_w)
        out_v.copy_(value.reshape(-1))


class BatchedMCTS:
    """Lockstep search over all trees of a :class:`TreePools`.

    ``evaluator(pools, out_w, out_v)`` must be stream-ordered GPU work that reads
    ``pools.leaf_planes`` (or ``leaf_me/leaf_opp``) and fills ``out_w`` float32 [B, A] (>= 0) and
    ``out_v`` float32 [B]; it is captured into the CUDA graph together with the tree kernel.
    """

    def __init__(self, pools: TreePools, evaluator, use_graph: bool = True, graph_unroll: int = 16, fused: bool = True):
        self.pools, self.evaluator = pools, evaluator
        self.use_graph, self.unroll, self.fused = bool(use_graph), int(graph_unroll), bool(fused)
        B, A = max(pools.n_trees, 1), pools.n_actions
        self.prior_w = torch.zeros((B, A), dtype=torch.float32, device=pools.device)
        self.value = torch.zeros(B, dtype=torch.float32, device=pools.device)
        self.counts = torch.zeros((B, A), dtype=torch.int32, device=pools.device)
        self.pi = torch.zeros((B, A), dtype=torch.float32, device=pools.device)
        self.q = torch.zeros((B, A), dtype=torch.float32, device=pools.device)
        self.best = torch.zeros(B, dtype=torch.uint8, device=pools.device)
        self._graph = None
        self._L = _lib.load()
        self.launches = 0  # kernels of libbetazero_b200 launched (bench.py's gpu_launches)

    # -- single kernels ------------------------------------------------------------------------
    def reset(self, root_me: torch.Tensor, root_opp: torch.Tensor) -> None:
        p = self.pools
        if root_me.numel() != p.n_trees or root_opp.numel() != p.n_trees:
            raise ValueError("one root per tree expected")
        _lib.check(self._L.bz_mcts_reset(p._ref, _lib.dptr(root_me), _lib.dptr(root_opp), _lib.stream_ptr()), "bz_mcts_reset")
        self.launches += 1

    def select(self) -> None:
        _lib.check(self._L.bz_mcts_select(self.pools._ref, _lib.stream_ptr()), "bz_mcts_select")
        self.launches += 1

    def evaluate(self) -> None:
        self.evaluator(self.pools, self.prior_w, self.value)

    def expand_backup(self) -> None:
        _lib.check(self._L.bz_mcts_expand_backup(self.pools._ref, _lib.dptr(self.prior_w), _lib.dptr(self.value),
                                                 _lib.stream_ptr()), "bz_mcts_expand_backup")
        self.launches += 1

    def step(self) -> None:
        if self.fused:
            _lib.check(self._L.bz_mcts_step(self.pools._ref, _lib.dptr(self.prior_w), _lib.dptr(self.value),
                                            _lib.stream_ptr()), "bz_mcts_step")
            self.launches += 1
        else:
            self.expand_backup()
            self.select()

    # -- the search ----------------------------------------------------------------------------
    def _capture(self) -> None:
        # warm up on a side stream (cuBLAS workspaces, lazy module init), then capture `unroll`
        # [evaluate -> step] blocks.  The warm-up really runs, so the pools must hold valid trees:
        # reset them to start positions first (the caller resets again for the real search).
        p = self.pools
        if p.game == GAME_TTT:
            rm = torch.zeros(p.n_trees, dtype=torch.int64, device=p.device)
            ro = torch.zeros_like(rm)
        else:
            from . import env

            rm, ro, _ = env.reversi_init(p.n_trees, p.board_size, device=p.device)
        self.reset(rm, ro)
        self.select()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self.evaluate()
                self.step()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        n0 = self.launches
        with torch.cuda.graph(g):
            for _ in range(self.unroll):
                self.evaluate()
                self.step()
        self._graph_launches = self.launches - n0
        self._graph = g

    def run(self, n_sims: int) -> None:
        """n_sims iterations on the current trees (after :meth:`reset`)."""
        if n_sims <= 0:
            return
        if n_sims > self.pools.sims_cap:
            raise ValueError(f"n_sims {n_sims} exceeds the pools' sims_cap {self.pools.sims_cap}")
        if self.pools.n_trees == 0:
            return
        self.select()
        inner = n_sims - 1
        if self.use_graph and inner >= self.unroll:
            if self._graph is None:
                raise RuntimeError("call prepare() before the first graph search")
            for _ in range(inner // self.unroll):
                self._graph.replay()
                self.launches += self._graph_launches
            inner %= self.unroll
        for _ in range(inner):
            self.evaluate()
            self.step()
        self.evaluate()
        self.expand_backup()

    def prepare(self) -> None:
        """Capture the CUDA graph (clobbers the pending leaf state: call before reset())."""
        if self.use_graph and self._graph is None and self.pools.n_trees > 0:
            self._capture()

    def search(self, root_me: torch.Tensor, root_opp: torch.Tensor, n_sims: int, check: bool = True):
        """Fresh search from the given roots.  Returns (visit_counts, pi, q) device tensors."""
        if self.use_graph and self._graph is None and n_sims - 1 >= self.unroll:
            self.prepare()
        self.reset(root_me, root_opp)
        self.run(n_sims)
        out = self.root_policy()
        if check:
            self.check_errors()
        return out

    def root_policy(self):
        p = self.pools
        _lib.check(self._L.bz_mcts_root_policy(p._ref, _lib.dptr(self.counts), _lib.dptr(self.pi), _lib.dptr(self.q),
                                               _lib.stream_ptr()), "bz_mcts_root_policy")
        self.launches += 1
        return self.counts, self.pi, self.q

    def best_action(self) -> torch.Tensor:
        _lib.check(self._L.bz_mcts_best_action(self.pools._ref, _lib.dptr(self.best), _lib.stream_ptr()),
                   "bz_mcts_best_action")
        self.launches += 1
        return self.best

    def check_errors(self) -> None:
        err = self.pools.error[: self.pools.n_trees]
        if self.pools.n_trees and bool(err.any().item()):
            codes = sorted(set(err[err != 0].tolist()))
            raise _lib.BzError(f"tree pool overflow (codes {codes}: 1 = edge pool, 2 = path depth); "
                               "enlarge edge_cap / max_depth")

    def stats(self) -> dict:
        """mean path depth d and mean edges per expanded node b of the last search (roofline model)."""
        p = self.pools
        sims = int(p.sim_count[: p.n_trees].sum().item())
        depth = int(p.depth_sum[: p.n_trees].sum().item())
        edges = int(p.edge_count[: p.n_trees].sum().item())
        return {"sims": sims, "mean_depth": depth / max(sims, 1), "edges": edges}
