"""CPU tests: the MCTS oracles.  The reference has no MCTS, so these pin (a) the C restatement
(oracle.c part 2) against the Python definition (oracle/mcts_ref.py), and (b) both against the
golden visit counts that mcts_ref.py produced when driving the LIVE reference board classes
(tests/golden/mcts.npz, oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle import mcts_ref as mr
from oracle import pyoracle as po
from oracle import ref_shim


def _c_search(me, opp, n_sims, salt, game, size, c_puct):
    return po.search_hash(np.array([me], np.uint64), np.array([opp], np.uint64), n_sims, game, size, c_puct, salt)


@pytest.mark.parametrize("prefix", ["rev8_playout_s48", "rev8_start_s400", "rev8_pass_s64"])
def test_c_mcts_matches_golden_reversi(golden_mcts, prefix):
    g = golden_mcts
    n_sims, salt = (int(v) for v in g[prefix + "_meta"])
    cnt, W, P, _ = po.search_hash(g[prefix + "_me"], g[prefix + "_opp"], n_sims, po.GAME_REVERSI, 8,
                                  float(g["c_puct"]), salt)
    assert np.array_equal(cnt, g[prefix + "_counts"])
    assert np.array_equal(W, g[prefix + "_W"]) and np.array_equal(P, g[prefix + "_P"])
    if prefix == "rev8_pass_s64":  # the only root action is pass, visited n_sims - 1 times
        assert (cnt[:, 64] == n_sims - 1).all() and (cnt[:, :64] == 0).all()


@pytest.mark.parametrize("n_sims", [25, 100])
def test_c_mcts_matches_golden_ttt_selfplay(golden_mcts, n_sims):
    """BASELINE config 1: tic-tac-toe MCTS self-play, root visit counts at every ply."""
    g = golden_mcts
    for salt in range(4):
        p = f"ttt_game_s{n_sims}_k{salt}"
        cnt, _, _, _ = po.search_hash(g[p + "_me"], g[p + "_opp"], n_sims, po.GAME_TTT, 3, float(g["c_puct"]), salt)
        assert np.array_equal(cnt, g[p + "_counts"])
        assert np.array_equal(cnt.argmax(1), g[p + "_action"])


@pytest.mark.parametrize("size,n_sims", [(4, 40), (6, 24)])
def test_c_mcts_matches_golden_small_boards(golden_mcts, size, n_sims):
    g = golden_mcts
    p = f"rev{size}_game_s{n_sims}"
    cnt, _, _, _ = po.search_hash(g[p + "_me"], g[p + "_opp"], n_sims, po.GAME_REVERSI, size, float(g["c_puct"]), 2)
    assert np.array_equal(cnt, g[p + "_counts"])


def test_python_definition_on_oracle_boards_matches_golden(golden_mcts):
    """mcts_ref.py driving the C-restated board classes reproduces what it produced on the live
    reference classes: the board restatement is interchangeable inside the search."""
    g = golden_mcts
    game = mr.ReversiGame(po.OracleReversiBoard, 8)
    n_sims, salt = (int(v) for v in g["rev8_playout_s48_meta"])
    for k in (0, 7, 23, 40, 58):
        me, opp = int(g["rev8_playout_s48_me"][k]), int(g["rev8_playout_s48_opp"][k])
        b = po.OracleReversiBoard(size=8)
        b.board = po.wire_to_grid(me, opp, 8)
        m = mr.MCTS(game, float(g["c_puct"]), lambda a, c: mr.hash_eval(a, c, salt, 65))
        m.reset(b, 1)
        m.run(n_sims)
        cnt, W, P = m.root_stats()
        assert np.array_equal(cnt, g["rev8_playout_s48_counts"][k])
        assert np.array_equal(W, g["rev8_playout_s48_W"][k]) and np.array_equal(P, g["rev8_playout_s48_P"][k])


def test_stepwise_c_tree_equals_python_definition_with_float_priors():
    """non-integer priors/values (like a real net's): C and Python must still agree bit for bit"""
    rng = np.random.default_rng(11)
    game = mr.ReversiGame(po.OracleReversiBoard, 8)
    table = {}

    def ev(me, opp):
        if (me, opp) not in table:
            w = rng.random(65).astype(np.float32) ** 3
            table[(me, opp)] = (w, np.float32(rng.uniform(-1, 1)))
        return table[(me, opp)]

    b, p = game.initial()
    m = mr.MCTS(game, 1.7, ev)
    m.reset(b, p)
    t = po.OracleTree(po.GAME_REVERSI, 8, 1.7)
    t.reset(np.asarray(b.board), p)
    for _ in range(300):
        st, me, opp = m.select()
        st2, me2, opp2, _ = t.select()
        assert (st, me, opp) == (st2, me2, opp2)
        w, v = ev(me, opp) if st == 0 else (None, 0.0)
        m.expand_backup(w, v)
        t.expand_backup(w, v)
    for a, b_ in zip(m.root_stats(), t.root_stats()):
        assert np.array_equal(a, b_)
    assert t.counters()["sum_depth"] == m.sum_depth


def test_hash_eval_c_equals_python():
    rng = np.random.default_rng(0)
    for _ in range(50):
        me, opp, salt = (int(v) for v in rng.integers(0, 2 ** 63, size=3))
        for A in (9, 65):
            w1, v1 = mr.hash_eval(me, opp, salt, A)
            w2, v2 = po.hash_eval(me, opp, salt, A)
            assert np.array_equal(w1, w2) and v1 == v2
            assert w1.min() >= 1 and w1.max() <= 32 and -1 <= v1 <= 0.875


def test_zero_weight_priors_fall_back_to_uniform():
    t = po.OracleTree(po.GAME_TTT, 3, 1.0)
    t.reset(np.zeros(9, np.int8), 1)
    t.select()
    t.expand_backup(np.zeros(9, np.float32), 0.0)
    _, _, P = t.root_stats()
    assert np.array_equal(P, np.full(9, np.float32(1) / np.float32(9), np.float32))


@pytest.mark.skipif(not ref_shim.available(), reason="live reference not present")
def test_definition_on_live_reference_matches_c(golden_mcts):
    RB = ref_shim.reversi_board_cls()
    game = mr.ReversiGame(RB, 8)
    b, p = game.initial()
    for (r, c) in ((2, 4), (2, 3)):
        b, p = game.next(b, p, r * 8 + c)
    m = mr.MCTS(game, 1.25, lambda a, c: mr.hash_eval(a, c, 9, 65))
    m.reset(b, p)
    m.run(120)
    me, opp = game.wire(b, p)
    cnt, W, P, _ = _c_search(me, opp, 120, 9, po.GAME_REVERSI, 8, 1.25)
    c0, w0, p0 = m.root_stats()
    assert np.array_equal(cnt[0], c0) and np.array_equal(W[0], w0) and np.array_equal(P[0], p0)


@pytest.mark.skipif(not ref_shim.available(), reason="live reference not present")
@pytest.mark.parametrize("n_sims", [25])
def test_reference_headless_loop_accepts_mcts_players(golden_mcts, n_sims):
    """BASELINE config 1 with the reference's OWN loop (TicTacToeHeadless.play, tic_tac_toe.py:13-34)
    and board class: an MCTS player with the reference's get_move interface reproduces the golden
    visit counts ply by ply (the GPU player does the same on the GPU box, tests/test_gpu_dropin.py)."""
    from helpers import OracleMCTSPlayer

    Headless = ref_shim.ttt_headless_cls()
    g = golden_mcts
    salt = 1
    counts = []

    class Rec(OracleMCTSPlayer):
        def get_move(self, board):
            mv = super().get_move(board)
            counts.append(self.last_counts.copy())
            return mv

    game = Headless(Rec(1, n_sims, game="ttt", salt=salt), Rec(-1, n_sims, game="ttt", salt=salt))
    positions, winner = game.play()
    p = f"ttt_game_s{n_sims}_k{salt}"
    assert np.array_equal(np.stack(counts), g[p + "_counts"]) and winner == int(g[p + "_winner"])
    assert len(positions) == len(counts) + 1


# ---- virtual loss: K descents per iteration (MCTS.select_vl / oracle.c Part 2d) ----------------------------------
VL_PREFIXES = ["rev8_playout_s48_k2", "rev8_playout_s48_k4", "rev8_start_s240_k4"]


@pytest.mark.parametrize("prefix", VL_PREFIXES)
def test_c_virtual_loss_matches_golden(golden_mcts_vl, prefix):
    """the C restatement reproduces what the Python definition computed on the LIVE reference boards"""
    g = golden_mcts_vl
    n_sims, salt, leaves = (int(v) for v in g[prefix + "_meta"])
    cnt, W, P, ctr = po.search_hash(g[prefix + "_me"], g[prefix + "_opp"], n_sims, po.GAME_REVERSI, 8,
                                    float(g["c_puct"]), salt, leaves=leaves)
    assert np.array_equal(cnt, g[prefix + "_counts"])
    assert np.array_equal(W, g[prefix + "_W"]) and np.array_equal(P, g[prefix + "_P"])
    assert ctr["sims"] == n_sims * len(cnt)


@pytest.mark.parametrize("leaves", [2, 4])
def test_c_virtual_loss_matches_golden_ttt_games(golden_mcts_vl, leaves):
    g = golden_mcts_vl
    p = f"ttt_game_s48_k{leaves}"
    cnt, _, _, _ = po.search_hash(g[p + "_me"], g[p + "_opp"], 48, po.GAME_TTT, 3, float(g["c_puct"]), 1, leaves=leaves)
    assert np.array_equal(cnt, g[p + "_counts"])
    assert np.array_equal(np.argmax(cnt, axis=1), g[p + "_action"])


def test_python_virtual_loss_definition_on_oracle_boards_matches_golden(golden_mcts_vl):
    """the definition driven by the C-restated board classes (what the GPU box can run) == driven by the reference"""
    g = golden_mcts_vl
    game = mr.ReversiGame(po.OracleReversiBoard, 8)
    prefix = "rev8_playout_s48_k4"
    n_sims, salt, leaves = (int(v) for v in g[prefix + "_meta"])
    for k in range(0, len(g[prefix + "_me"]), 5):
        board = po.OracleReversiBoard(size=8)
        board.board = po.wire_to_grid(g[prefix + "_me"][k], g[prefix + "_opp"][k], 8)
        m = mr.MCTS(game, float(g["c_puct"]), lambda a, b: mr.hash_eval(a, b, salt, 65))
        m.reset(board, 1)
        m.run_vl(n_sims, leaves)
        c, w, p_ = m.root_stats()
        assert np.array_equal(c, g[prefix + "_counts"][k]) and np.array_equal(w, g[prefix + "_W"][k])
        assert np.array_equal(p_, g[prefix + "_P"][k])


def test_virtual_loss_with_one_leaf_is_the_sequential_search():
    """K = 1: 'descents that entered the node before' == 1 + sum(child N): same visit counts and priors as MCTS.run;
    W differs only by the (W - 1) + 1 roundings"""
    me, opp = po.playout_boards(24, seed=2)
    c1, W1, P1, k1 = po.search_hash(me, opp, 200, salt=4)
    trees = [po.OracleTree() for _ in range(len(me))]
    for t, m, o in zip(trees, me, opp):
        t.reset_wire(m, o)
        for _ in range(200):
            st, lm, lo, _ = t.select_vl(0)
            w, v = po.hash_eval(lm, lo, 4, 65) if st == 0 else (None, 0.0)
            t.expand_backup_vl(0, w, v)
    for i, t in enumerate(trees):
        c, W, P = t.root_stats()
        assert np.array_equal(c, c1[i]) and np.array_equal(P, P1[i])
        np.testing.assert_allclose(W, W1[i], atol=2e-5)


def test_virtual_loss_stepwise_c_equals_python_with_float_priors():
    """random float (w, v) per leaf, 3 leaves per iteration: leaves, statuses, depths and statistics agree"""
    game = mr.ReversiGame(po.OracleReversiBoard, 8)
    b, p = game.initial()
    rng = np.random.default_rng(5)
    m = mr.MCTS(game, 2.0, None)
    m.reset(b, p)
    t = po.OracleTree(po.GAME_REVERSI, 8, 2.0)
    t.reset(np.asarray(b.board), p)
    K = 3
    for it in range(70):
        evs = []
        for j in range(K):
            st, me, opp = m.select_vl(j)
            cst, cme, copp, cd = t.select_vl(j)
            assert (st, me, opp) == (cst, cme, copp), (it, j)
            w = (rng.random(65) ** 3).astype(np.float32)
            w[rng.random(65) < 0.1] = 0
            evs.append((w, np.float32(rng.uniform(-1, 1))))
        for j in range(K):
            m.expand_backup_vl(j, *evs[j])
            t.expand_backup_vl(j, *evs[j])
    for a, b_ in zip(m.root_stats(), t.root_stats()):
        assert np.array_equal(a, b_)
    assert t.counters()["sims"] == 70 * K


@pytest.mark.skipif(not ref_shim.available(), reason="live reference not present")
def test_virtual_loss_definition_on_live_reference_matches_c():
    RB = ref_shim.reversi_board_cls()
    game = mr.ReversiGame(RB, 8)
    b, p = game.initial()
    for (r, c) in ((2, 4), (2, 3)):
        b, p = game.next(b, p, r * 8 + c)
    m = mr.MCTS(game, 1.25, lambda a, c: mr.hash_eval(a, c, 9, 65))
    m.reset(b, p)
    m.run_vl(120, 4)
    me, opp = game.wire(b, p)
    cnt, W, P, _ = po.search_hash(np.array([me], np.uint64), np.array([opp], np.uint64), 120, po.GAME_REVERSI, 8, 1.25, 9,
                                  leaves=4)
    c0, w0, p0 = m.root_stats()
    assert np.array_equal(cnt[0], c0) and np.array_equal(W[0], w0) and np.array_equal(P[0], p0)


@pytest.mark.parametrize("leaves", [1, 4])
@pytest.mark.parametrize("cap", [None, 120])
def test_tree_reuse_c_restatement_matches_the_python_definition(leaves, cap):
    """MCTS.advance (keep the subtree of the move played iff expanded and within the cap; the next search adds to its
    statistics) against orc_mcts_advance / orc_mcts_play_hash: same visit counts and moves at every ply of a game."""
    game = mr.ReversiGame(po.OracleReversiBoard, 6)
    salt = 5
    hist, _ = mr.self_play_game(game, 32, 1.25, lambda a, b: mr.hash_eval(a, b, salt, 65), max_plies=14, leaves=leaves,
                                reuse=True, cap_units=cap)
    b, p = game.initial()
    me, opp = game.wire(b, p)
    cnt, act, kept = po.play_hash([me], [opp], 32, 14, size=6, salt=salt, leaves=leaves, reuse=True,
                                  cap_units=-1 if cap is None else cap)
    assert len(hist) == 14
    for i, h in enumerate(hist):
        assert np.array_equal(h[3], cnt[0, i]) and h[4] == act[0, i], i
    assert kept[0].sum() >= 6  # most moves continue on a kept subtree ...
    assert any(int(c.sum()) > 32 for c in cnt[0])  # ... whose root carries more than one search's visits
    if cap is not None:
        assert kept[0].sum() < 14  # the cap drops the large ones
    # without reuse the same driver reproduces the fresh-tree games
    hist0, _ = mr.self_play_game(game, 32, 1.25, lambda a, b: mr.hash_eval(a, b, salt, 65), max_plies=6, leaves=leaves)
    cnt0, act0, kept0 = po.play_hash([me], [opp], 32, 6, size=6, salt=salt, leaves=leaves)
    assert all(np.array_equal(h[3], cnt0[0, i]) and h[4] == act0[0, i] for i, h in enumerate(hist0)) and not kept0.any()
