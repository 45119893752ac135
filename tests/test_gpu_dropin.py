"""GPU: the drop-in boundary.  The reference's own callers -- its episode loops and player
interface -- run on the CUDA-backed board / player classes and reproduce the oracle exactly."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from helpers import OracleMCTSPlayer, play_reversi_episode, play_ttt_headless  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

pytestmark = pytest.mark.gpu


class ScriptedPlayer:
    """plays the k-th legal move (row-major), k from a seeded stream: same choices for any board class"""

    def __init__(self, symbol, seed):
        self.symbol, self.rng = symbol, np.random.default_rng(seed)

    def get_move(self, board):
        moves = board.generate_possible_moves(self.symbol)
        return moves[int(self.rng.integers(len(moves)))]


@pytest.mark.parametrize("size", [4, 6, 8])
def test_reversi_board_api_matches_oracle_through_reference_loop(size):
    from betazero_b200.boards import ReversiBoard

    gb, gtrace = play_reversi_episode(ReversiBoard, ScriptedPlayer(1, 5), ScriptedPlayer(-1, 6), size)
    ob, otrace = play_reversi_episode(po.OracleReversiBoard, ScriptedPlayer(1, 5), ScriptedPlayer(-1, 6), size)
    assert gtrace == otrace and np.array_equal(gb.board, ob.board)
    assert gb.is_game_over() and gb.get_score() == ob.get_score()
    # method by method along the way
    g, o = ReversiBoard(size=size), po.OracleReversiBoard(size=size)
    assert np.array_equal(g.board, o.board) and g.board.dtype.kind == "i"
    for (pl, mv) in otrace[: 12 if size == 8 else None]:
        assert g.generate_possible_moves(1) == o.generate_possible_moves(1)
        assert g.generate_possible_moves(-1) == o.generate_possible_moves(-1)
        assert g.is_game_over() == o.is_game_over() and g.get_score() == o.get_score()
        for r in range(size):
            for c in range(size):
                assert g.is_valid_move(r, c, pl) == o.is_valid_move(r, c, pl)
        if mv is not None:
            g2 = g.make_move(*mv, pl)
            assert np.array_equal(g.board, o.board)  # value semantics: make_move never mutates self
            g, o = g2, o.make_move(*mv, pl)
            assert np.array_equal(g.board, o.board)
    cp = ReversiBoard(g)  # copy-ctor takes a board object (reversi_board.py:13-14)
    assert np.array_equal(cp.board, g.board) and cp.size == g.size and cp.board is not g.board
    assert not g.is_valid_move(-1, 0, 1) and not g.is_valid_move(0, size, 1)


def test_reversi_invalid_move_raises_value_error():
    from betazero_b200.boards import ReversiBoard

    b = ReversiBoard(size=8)
    for (r, c) in ((0, 0), (3, 3), (9, 9), (-1, 2)):
        with pytest.raises(ValueError, match="Invalid move"):
            b.make_move(r, c, 1)
    assert b.make_move(2, 4, 1).board[3, 4] == 1  # a legal opening move flips (3,4)


def test_reference_demo_sequence(golden_env):
    """reversi_board.py:92-99 on 4x4, expected boards recorded from the reference"""
    from betazero_b200.boards import ReversiBoard

    b = ReversiBoard(size=4)
    demo = golden_env["demo4_boards"]
    assert np.array_equal(b.board, demo[0])
    for k, (r, c, p) in enumerate(((0, 2, 1), (0, 1, -1), (2, 0, 1))):
        b = b.make_move(r, c, p)
        assert np.array_equal(b.board, demo[k + 1])
    assert b.is_valid_move(0, 3, -1) == bool(golden_env["demo4_valid_0_3_m1"])
    assert "X" in str(b) and repr(b) == f"{b.board}"


def test_ttt_board_api_matches_oracle():
    from betazero_b200.boards import TicTacToeBoard

    rng = np.random.default_rng(2)
    for _ in range(40):
        g = rng.integers(-1, 2, size=(3, 3))
        gb, ob = TicTacToeBoard(g), po.OracleTicTacToeBoard(g)
        assert gb.is_game_over() == ob.is_game_over()
        assert gb.generate_possible_moves() == ob.generate_possible_moves()
        for r in range(3):
            for c in range(3):
                assert gb.is_valid_move(r, c) == ob.is_valid_move(r, c)
                if ob.is_valid_move(r, c):
                    assert np.array_equal(gb.make_move(r, c, -1).board, ob.make_move(r, c, -1).board)
                else:
                    with pytest.raises(ValueError, match="Invalid move"):
                        gb.make_move(r, c, 1)
    assert TicTacToeBoard().is_game_over() == (False, None)


@pytest.mark.parametrize("n_sims", [25, 100])
def test_config1_ttt_mcts_selfplay_through_reference_loop(golden_mcts, n_sims):
    """BASELINE config 1: tic-tac-toe MCTS self-play through the reference's headless loop; the
    root visit counts of every ply equal the golden ones (mcts_ref.py on the live reference)."""
    from betazero_b200.boards import TicTacToeBoard
    from betazero_b200.players import TicTacToeMCTSPlayer

    g = golden_mcts
    for salt in range(2):
        p = f"ttt_game_s{n_sims}_k{salt}"
        counts = []

        class Rec(TicTacToeMCTSPlayer):
            def get_move(self, board):
                mv = super().get_move(board)
                counts.append(self.last_counts.copy())
                return mv

        p1, p2 = Rec(1, n_sims=n_sims, salt=salt), Rec(-1, n_sims=n_sims, salt=salt)
        positions, winner = play_ttt_headless(TicTacToeBoard, p1, p2)
        assert np.array_equal(np.stack(counts), g[p + "_counts"])
        assert winner == int(g[p + "_winner"])
        assert len(positions) == len(counts) + 1
        # positions are the raw (un-canonicalised) boards, as TicTacToeHeadless records them
        for i, pos in enumerate(positions[:-1]):
            pl = 1 if i % 2 == 0 else -1
            me = sum(1 << k for k in range(9) if pos.reshape(-1)[k] == pl)
            assert me == int(g[p + "_me"][i])


@pytest.mark.parametrize("size,n_sims", [(4, 40), (6, 24)])
def test_reversi_mcts_player_through_reference_loop(golden_mcts, size, n_sims):
    from betazero_b200.boards import ReversiBoard
    from betazero_b200.players import MCTSPlayer

    g = golden_mcts
    p = f"rev{size}_game_s{n_sims}"
    p1, p2 = MCTSPlayer(1, n_sims=n_sims, size=size, salt=2), MCTSPlayer(-1, n_sims=n_sims, size=size, salt=2)
    board, trace = play_reversi_episode(ReversiBoard, p1, p2, size)
    exp = [None if a == 64 else (int(a) >> 3, int(a) & 7) for a in g[p + "_action"]]
    assert [mv for _, mv in trace] == exp
    assert [pl for pl, _ in trace] == [int(v) for v in g[p + "_player"]]
    assert board.is_game_over() and board.get_score()[0] == int(g[p + "_winner"])


def test_mcts_player_vs_oracle_player_8x8_opening():
    from betazero_b200.boards import ReversiBoard
    from betazero_b200.players import MCTSPlayer

    gp, op = MCTSPlayer(1, n_sims=60, salt=4), OracleMCTSPlayer(1, 60, salt=4)
    b = ReversiBoard(size=8)
    for _ in range(3):
        mv = gp.get_move(b)
        assert mv == op.get_move(b) and np.array_equal(gp.last_counts, op.last_counts)
        b = b.make_move(*mv, 1)
        b = b.make_move(*b.generate_possible_moves(-1)[0], -1)


def test_greedy_net_player_picks_best_legal_move():
    from betazero_b200 import net
    from betazero_b200.boards import ReversiBoard
    from betazero_b200.players import GreedyNetPlayer

    b = ReversiBoard(size=8)
    pl = GreedyNetPlayer(1, net.make_net("mlp", seed=1))
    mv = pl.get_move(b)
    assert mv in b.generate_possible_moves(1)


@pytest.mark.parametrize("leaves", [2, 4])
def test_ttt_mcts_players_with_virtual_loss_through_reference_loop(golden_mcts_vl, leaves):
    """the drop-in player with n_leaves > 1 (single-tree wave kernels) in the reference's headless loop: per-ply root
    visit counts and the winner equal the virtual-loss golden game (mcts_ref.MCTS.run_vl on the live reference)"""
    from betazero_b200.boards import TicTacToeBoard
    from betazero_b200.players import TicTacToeMCTSPlayer

    g = golden_mcts_vl
    p = f"ttt_game_s48_k{leaves}"
    counts = []

    class Rec(TicTacToeMCTSPlayer):
        def get_move(self, board):
            mv = super().get_move(board)
            counts.append(self.last_counts.copy())
            return mv

    positions, winner = play_ttt_headless(TicTacToeBoard, Rec(1, n_sims=48, salt=1, n_leaves=leaves),
                                          Rec(-1, n_sims=48, salt=1, n_leaves=leaves))
    assert np.array_equal(np.stack(counts), g[p + "_counts"])
    assert winner == int(g[p + "_winner"])
