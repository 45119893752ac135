"""CPU tests: pin the C oracle (oracle/oracle.c) against outputs of the reference itself.

tests/golden/reversi_env.npz and ttt_env.npz were produced by oracle/make_golden.py running the
LIVE reference (reversi_board.py / tic_tac_toe_board.py); ttt_env.npz also carries the
reference's only golden file, tic_tac_toe_data.csv.  When /root/reference is present the oracle is
additionally checked against the live classes on fresh random boards.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from oracle import ref_shim


@pytest.mark.parametrize("size", [4, 6, 8])
def test_oracle_matches_reference_outputs(golden_env, size):
    g = golden_env
    me, opp = g[f"s{size}_me"], g[f"s{size}_opp"]
    assert np.array_equal(po.legal_mask(me, opp, size), g[f"s{size}_mask"])
    assert np.array_equal(po.legal_mask(opp, me, size), g[f"s{size}_mask_opp"])
    over, win, cm, co = po.terminal(me, opp, size)
    assert np.array_equal(over, g[f"s{size}_over"])
    assert np.array_equal(win, g[f"s{size}_winner"])
    assert np.array_equal(cm, g[f"s{size}_cnt_me"]) and np.array_equal(co, g[f"s{size}_cnt_opp"])
    idx, act = g[f"s{size}_succ_idx"], g[f"s{size}_succ_act"]
    mo, oo, err = po.apply(me[idx], opp[idx], act, size)
    assert not err.any()
    assert np.array_equal(mo, g[f"s{size}_succ_me"]) and np.array_equal(oo, g[f"s{size}_succ_opp"])
    bidx, bact = g[f"s{size}_bad_idx"], g[f"s{size}_bad_act"]
    _, _, err = po.apply(me[bidx], opp[bidx], bact, size)
    assert err.all()  # the reference raised ValueError("Invalid move") on each of these


def test_oracle_board_class_replays_reference_demo(golden_env):
    # reversi_board.py:92-99 demo on 4x4; expected boards recorded from the reference
    b = po.OracleReversiBoard(size=4)
    demo = golden_env["demo4_boards"]
    assert np.array_equal(b.board, demo[0])
    for k, (r, c, p) in enumerate(((0, 2, 1), (0, 1, -1), (2, 0, 1))):
        b = b.make_move(r, c, p)
        assert np.array_equal(b.board, demo[k + 1])
    assert b.is_valid_move(0, 3, -1) == bool(golden_env["demo4_valid_0_3_m1"])
    with pytest.raises(ValueError, match="Invalid move"):
        b.make_move(0, 2, 1)  # occupied cell


@pytest.mark.parametrize("size", [4, 6, 8])
def test_oracle_episode_matches_reference(golden_env, size):
    # full reference episode (reversi_terminal.py:16-38 order): replay its actions through the oracle
    g = golden_env
    me, opp, act = g[f"ep{size}_me"], g[f"ep{size}_opp"], g[f"ep{size}_action"]
    for k in range(len(act) - 1):
        m, o, err = po.apply(me[k:k + 1], opp[k:k + 1], act[k:k + 1], size)
        assert err[0] == 0 and m[0] == me[k + 1] and o[0] == opp[k + 1]
    m, o, err = po.apply(me[-1:], opp[-1:], act[-1:], size)
    over, win, cm, co = po.terminal(m, o, size)
    assert over[0] == 1
    final = g[f"ep{size}_final"]
    w, c1, c2 = g[f"ep{size}_score"]
    b = po.OracleReversiBoard(size=size)
    b.board = final.astype(np.int64)
    assert b.is_game_over() and b.get_score() == (w, (c1, c2))


def test_ttt_oracle_matches_reference_outputs(golden_ttt):
    g = golden_ttt
    assert np.array_equal(po.ttt_legal_mask(g["x"], g["o"]), g["mask"])
    over, win = po.ttt_terminal(g["x"], g["o"])
    assert np.array_equal(over, g["over"]) and np.array_equal(win, g["winner"])


def test_ttt_csv_golden(golden_ttt):
    """tic_tac_toe_data.csv: canonical states (mover = +1) and one-hot actions of optimal play.
    Pins canonical form (generate_training_games.py:17-21) and move application."""
    st, ac = golden_ttt["csv_state"], golden_ttt["csv_action"]
    assert st.shape == (180, 9) and ac.shape == (180, 9)
    for s, a in zip(st, ac):
        assert a.sum() == 1 and set(np.unique(a)) <= {0, 1}
        k = int(np.argmax(a))
        # canonical: the mover is +1, so it has as many marks as the opponent, or one fewer
        assert (s == -1).sum() - (s == 1).sum() in (0, 1)
        b = po.OracleTicTacToeBoard(s.reshape(3, 3).astype(np.int64))
        assert b.is_game_over() == (False, None)
        assert (k // 3, k % 3) in b.generate_possible_moves()
        nb = b.make_move(k // 3, k % 3, 1)
        assert nb.board[k // 3, k % 3] == 1
        # wire-format equivalents
        x = sum(1 << i for i in range(9) if s[i] == 1)
        o = sum(1 << i for i in range(9) if s[i] == -1)
        xo, oo, err = po.ttt_apply([x], [o], [k], [1])
        assert err[0] == 0 and xo[0] == x | (1 << k) and oo[0] == o
    # consecutive rows of one game: next state == -(state + action) (generate_training_games.py:21)
    follows = sum(np.array_equal(st[i + 1], -(st[i] + ac[i])) for i in range(179))
    assert follows >= 180 - 20 - 20  # 20 games: every in-game transition must follow the rule


@pytest.mark.skipif(not ref_shim.available(), reason="live reference not present")
def test_oracle_vs_live_reference_random_boards():
    RB = ref_shim.reversi_board_cls()
    rng = np.random.default_rng(123)
    for size in (4, 6, 8):
        for _ in range(150):
            pe = rng.uniform(0.05, 0.9)
            u = rng.random((size, size))
            g = np.where(u < pe, 0, np.where(rng.random((size, size)) < 0.5, 1, -1))
            rb = RB(size=size)
            rb.board = g.copy()
            ob = po.OracleReversiBoard(size=size)
            ob.board = g.copy()
            for p in (1, -1):
                assert rb.generate_possible_moves(p) == ob.generate_possible_moves(p)
                for (r, c) in rb.generate_possible_moves(p)[:4]:
                    assert np.array_equal(rb.make_move(r, c, p).board, ob.make_move(r, c, p).board)
            assert rb.is_game_over() == ob.is_game_over()
            assert rb.get_score() == ob.get_score()


@pytest.mark.skipif(not ref_shim.available(), reason="live reference not present")
def test_ttt_oracle_vs_live_reference():
    TB = ref_shim.ttt_board_cls()
    rng = np.random.default_rng(5)
    for _ in range(300):
        g = rng.integers(-1, 2, size=(3, 3))
        rb, ob = TB(g), po.OracleTicTacToeBoard(g)
        assert rb.is_game_over() == ob.is_game_over()
        assert rb.generate_possible_moves() == ob.generate_possible_moves()


def test_board_generators_are_seeded():
    a = po.synthetic_boards(1000, seed=0)
    b = po.synthetic_boards(1000, seed=0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert not (a[0] & a[1]).any()
    m, o = po.playout_boards(64, seed=0)
    assert not (m & o).any()
    # reachable boards keep the four centre cells occupied
    centre = np.uint64((1 << 27) | (1 << 28) | (1 << 35) | (1 << 36))
    assert (((m | o) & centre) == centre).all()
