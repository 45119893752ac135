"""Batched lockstep self-play: the reference episode loop (reversi_terminal.py:16-38) for
thousands of concurrent games per GPU.

Per ply, for ALL game slots at once:
    search (BatchedMCTS: n_sims iterations) -> bz_selfplay_advance (pick move from root visit
    counts, record (board, pi), apply / pass, terminal test, score + flush finished games to the
    replay buffer, restart the slot) -> next search from the new positions.
Games shard embarrassingly: rank r of W owns game ids {r*B + s + k*W*B}; move sampling is keyed by
(seed, game id, ply), so a game's content does not depend on which GPU / slot played it.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import BzSelfplayState
from .mcts import BatchedMCTS, TreePools

N_ACTIONS = 65


class BatchedSelfPlay:
    def __init__(self, n_games: int, n_sims: int, evaluator, board_size: int = 8, c_puct: float = 1.25,
                 temp_plies: int = 0, seed: int = 0, replay_cap: int | None = None, rank: int = 0, world: int = 1,
                 arena_units: int | None = None, use_graph: bool = True, graph_unroll: int = 16,
                 dirichlet_alpha: float = 0.0, dirichlet_eps: float = 0.25, n_leaves: int = 1, device="cuda",
                 one_launch: bool | None = None, reuse: bool = False):
        self.n_games, self.n_sims, self.board_size = int(n_games), int(n_sims), int(board_size)
        self.rank, self.world = int(rank), int(world)
        self.device = torch.device(device)
        # reuse: a move's search continues on the subtree of the move played (opt-in; bz_mcts_reroot) -- n_sims MORE
        # simulations per move on top of the kept ones; the default is a fresh tree per move, which the goldens pin
        self.reuse = bool(reuse)
        self._moves = 0
        self.pools = TreePools(n_games, n_sims, board_size=board_size, c_puct=c_puct, arena_units=arena_units,
                               n_leaves=n_leaves, device=device, reuse=reuse)
        self.mcts = BatchedMCTS(self.pools, evaluator, use_graph=use_graph, graph_unroll=graph_unroll,
                                dirichlet_alpha=dirichlet_alpha, dirichlet_eps=dirichlet_eps,
                                noise_seed=seed * 7919 + rank, one_launch=one_launch)
        self.max_plies = 128
        B, H = max(self.n_games, 1), max(self.n_games, 1) * self.max_plies
        self.replay_cap = int(replay_cap if replay_cap is not None else B * 2 * self.max_plies)
        R = max(self.replay_cap, 1)
        dev = self.device

        def e(n, dt):
            return torch.empty(n, dtype=dt, device=dev)

        self.me, self.opp = e(B, torch.int64), e(B, torch.int64)
        self.player, self.ply, self.game_id = e(B, torch.int8), e(B, torch.int32), e(B, torch.int64)
        self.hist_me, self.hist_opp = e(H, torch.int64), e(H, torch.int64)
        self.hist_player, self.hist_action = e(H, torch.int8), e(H, torch.uint8)
        self.hist_pi = e(H * N_ACTIONS, torch.float32)
        self.rp_me, self.rp_opp = e(R, torch.int64), e(R, torch.int64)
        self.rp_pi = e(R * N_ACTIONS, torch.float32)
        self.rp_z, self.rp_game, self.rp_ply = e(R, torch.int8), e(R, torch.int64), e(R, torch.int16)
        self.counters = torch.zeros(8, dtype=torch.int64, device=dev)
        self.last_action = torch.zeros(B, dtype=torch.uint8, device=dev)
        s = BzSelfplayState()
        s.n_games, s.board_size, s.max_plies, s.temp_plies = self.n_games, self.board_size, self.max_plies, int(temp_plies)
        s.seed, s.id_stride, s.replay_cap = int(seed), self.world * self.n_games, self.replay_cap
        for name, _ in BzSelfplayState._fields_[7:]:
            setattr(s, name, getattr(self, name).data_ptr())
        self.c_struct = s
        self._ref = C.byref(s)
        self._L = _lib.load()
        self.launches = 0
        self._dropped_seen = 0  # counters[5] at the last drain (overflow is reported, never silent)
        _lib.check(self._L.bz_selfplay_init(self._ref, self.rank * self.n_games, _lib.stream_ptr()), "bz_selfplay_init")
        self.launches += 1

    # -- one lockstep ply for every game ----------------------------------------------------------
    def search(self) -> None:
        if (self.mcts.use_graph and self.mcts._graph is None and not self.mcts.one_launch
                and self.n_sims // self.pools.n_leaves - 1 >= self.mcts.unroll):
            self.mcts.prepare()  # lazily, before the roots are set: play_move() works without an explicit prepare()
        if self.reuse and self._moves > 0:
            self.mcts.advance(self.last_action, self.me, self.opp)  # slots whose game restarted get an empty tree
        else:
            self.mcts.reset(self.me, self.opp)
        self.mcts.run(self.n_sims)
        self._moves += 1

    def advance(self) -> None:
        _lib.check(self._L.bz_selfplay_advance(self._ref, self.pools._ref, _lib.dptr(self.last_action),
                                               _lib.stream_ptr()), "bz_selfplay_advance")
        self.launches += 1

    def play_move(self) -> None:
        self.search()
        self.advance()

    def prepare(self) -> None:
        """capture the CUDA graph (before the first search)"""
        self.mcts.prepare()

    # -- results ---------------------------------------------------------------------------------
    def stats(self) -> dict:
        c = self.counters.cpu().tolist()
        return {"replay_records": c[0], "plies": c[1], "o_wins": c[2], "draws": c[3], "x_wins": c[4],
                "dropped": c[5], "games": c[6]}

    def drain_replay(self) -> dict:
        """Returns the finished games' records (device tensors, copies) and empties the buffer."""
        c = self.counters.cpu().tolist()
        n = min(int(c[0]), self.replay_cap)  # counters[0] counts committed rows only (a game that does not fit takes none)
        if c[5] > self._dropped_seen:
            import warnings

            warnings.warn(f"replay buffer overflow: {c[5] - self._dropped_seen} records of finished games were dropped "
                          f"(replay_cap {self.replay_cap}); drain more often or enlarge replay_cap", RuntimeWarning)
            self._dropped_seen = c[5]
        out = {
            "me": self.rp_me[:n].clone(), "opp": self.rp_opp[:n].clone(),
            "pi": self.rp_pi[: n * N_ACTIONS].view(n, N_ACTIONS).clone(),
            "z": self.rp_z[:n].clone(), "game": self.rp_game[:n].clone(), "ply": self.rp_ply[:n].clone(),
        }
        self.counters[0] = 0
        return out

    def total_launches(self) -> int:
        return self.launches + self.mcts.launches
