"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the LIVE reference.

Run in the authoring container (needs /root/reference; nothing is copied from it):

    python oracle/make_golden.py

Every array below is an OUTPUT of the reference's own code
(src/reversi/game_logic/reversi_board.py, src/tic_tac_toe/tic_tac_toe_board.py) -- or, for the
MCTS fixtures, of oracle/mcts_ref.py driving those reference classes -- on seeded inputs.
The fixtures pin (a) the C oracle on the GPU box, where the reference does not exist, and
(b) the CUDA kernels directly.

Wire format: Reversi boards are mover-relative (me, opp) uint64 pairs, bit = row*8 + col for
every board size; the mover is always the reference's player +1 in these fixtures.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mcts_ref as mr  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def wire(grid, player=1):
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


def grid_from_wire(me, opp, size):
    g = np.zeros((size, size), dtype=int)
    for r in range(size):
        for c in range(size):
            b = r * 8 + c
            if (me >> b) & 1:
                g[r, c] = 1
            elif (opp >> b) & 1:
                g[r, c] = -1
    return g


def synthetic(n, size, rng):
    """iid cells, p_empty ~ U[0.05, 0.9] per board (SURVEY.md 8d set A)."""
    out = []
    for _ in range(n):
        pe = rng.uniform(0.05, 0.9)
        u = rng.random((size, size))
        s = rng.random((size, size)) < 0.5
        g = np.where(u < pe, 0, np.where(s, 1, -1))
        out.append(g)
    return out


def playouts(n, size, rng, RB):
    """Boards reached by the reference loop (reversi_terminal.py:16-38) with uniformly random
    movers, stopped after a random number of loop iterations; always re-expressed with the side
    to move as +1 (canonical form, players.py:85)."""
    out = []
    while len(out) < n:
        b, cur, over = RB(size=size), 1, False
        stop = int(rng.integers(0, size * size))
        it = 0
        while not over and it < stop:
            moves = b.generate_possible_moves(cur)
            if moves:
                r, c = moves[int(rng.integers(len(moves)))]
                b = b.make_move(r, c, cur)
            over = b.is_game_over()
            cur *= -1
            it += 1
        out.append(np.asarray(b.board) * cur)
    return out


def reversi_env(RB):
    rng = np.random.default_rng(20261018)
    recs = {}
    for size, n_syn, n_play in ((8, 3000, 1500), (6, 400, 300), (4, 300, 200)):
        grids = synthetic(n_syn, size, rng) + playouts(n_play, size, rng, RB)
        # hand-made edge cases: empty, full, start position, single colour
        grids += [np.zeros((size, size), int), np.ones((size, size), int), -np.ones((size, size), int),
                  np.asarray(RB(size=size).board), -np.asarray(RB(size=size).board)]
        n = len(grids)
        me = np.zeros(n, np.uint64)
        opp = np.zeros(n, np.uint64)
        mask = np.zeros(n, np.uint64)
        mask_opp = np.zeros(n, np.uint64)
        over = np.zeros(n, np.uint8)
        winner = np.zeros(n, np.int8)
        c1 = np.zeros(n, np.uint8)
        c2 = np.zeros(n, np.uint8)
        succ_idx, succ_act, succ_me, succ_opp = [], [], [], []
        bad_idx, bad_act = [], []
        for i, g in enumerate(grids):
            b = RB(size=size)
            b.board = np.array(g, dtype=int)
            m, o = wire(g, 1)
            me[i], opp[i] = m, o
            mv = b.generate_possible_moves(1)
            mask[i] = sum(1 << (r * 8 + c) for r, c in mv)
            mask_opp[i] = sum(1 << (r * 8 + c) for r, c in b.generate_possible_moves(-1))
            over[i] = b.is_game_over()
            w, (a, bb) = b.get_score()
            winner[i], c1[i], c2[i] = w, a, bb
            for r, c in mv:  # every legal successor, re-expressed for the next mover (-1)
                nb = b.make_move(r, c, 1)
                sm, so = wire(nb.board, -1)
                succ_idx.append(i)
                succ_act.append(r * 8 + c)
                succ_me.append(sm)
                succ_opp.append(so)
            # a few illegal (row, col) per board must raise ValueError("Invalid move")
            for _ in range(2):
                r, c = int(rng.integers(size)), int(rng.integers(size))
                if (r, c) not in mv:
                    try:
                        b.make_move(r, c, 1)
                        raise AssertionError("reference accepted an illegal move")
                    except ValueError:
                        bad_idx.append(i)
                        bad_act.append(r * 8 + c)
        recs[f"s{size}_me"], recs[f"s{size}_opp"] = me, opp
        recs[f"s{size}_mask"], recs[f"s{size}_mask_opp"] = mask, mask_opp
        recs[f"s{size}_over"], recs[f"s{size}_winner"] = over, winner
        recs[f"s{size}_cnt_me"], recs[f"s{size}_cnt_opp"] = c1, c2
        recs[f"s{size}_succ_idx"] = np.array(succ_idx, np.int32)
        recs[f"s{size}_succ_act"] = np.array(succ_act, np.uint8)
        recs[f"s{size}_succ_me"] = np.array(succ_me, np.uint64)
        recs[f"s{size}_succ_opp"] = np.array(succ_opp, np.uint64)
        recs[f"s{size}_bad_idx"] = np.array(bad_idx, np.int32)
        recs[f"s{size}_bad_act"] = np.array(bad_act, np.uint8)
        print(f"reversi size {size}: {n} boards, {len(succ_idx)} successors, {len(bad_idx)} illegal moves")
    # the reference demo sequence (reversi_board.py:92-99) on 4x4, outputs recorded here
    b = RB(size=4)
    demo = [np.asarray(b.board).copy()]
    for (r, c, p) in ((0, 2, 1), (0, 1, -1), (2, 0, 1)):
        b = b.make_move(r, c, p)
        demo.append(np.asarray(b.board).copy())
    recs["demo4_boards"] = np.stack(demo).astype(np.int8)
    recs["demo4_valid_0_3_m1"] = np.array(b.is_valid_move(0, 3, -1))
    # one full reference episode per size (reversi_terminal.py loop, first-legal-move players)
    for size in (4, 6, 8):
        b, cur, over = RB(size=size), 1, False
        tr_me, tr_opp, tr_pl, tr_act = [], [], [], []
        while not over:
            moves = b.generate_possible_moves(cur)
            m, o = wire(b.board, cur)
            tr_me.append(m)
            tr_opp.append(o)
            tr_pl.append(cur)
            if moves:
                r, c = moves[len(moves) // 2]
                tr_act.append(r * 8 + c)
                b = b.make_move(r, c, cur)
            else:
                tr_act.append(64)
            over = b.is_game_over()
            cur *= -1
        w, (a, bb) = b.get_score()
        recs[f"ep{size}_me"] = np.array(tr_me, np.uint64)
        recs[f"ep{size}_opp"] = np.array(tr_opp, np.uint64)
        recs[f"ep{size}_player"] = np.array(tr_pl, np.int8)
        recs[f"ep{size}_action"] = np.array(tr_act, np.uint8)
        recs[f"ep{size}_final"] = np.asarray(b.board).astype(np.int8)
        recs[f"ep{size}_score"] = np.array([w, a, bb], np.int32)
    np.savez_compressed(os.path.join(OUT, "reversi_env.npz"), **recs)


def ttt_env(TB):
    # all 3^9 grids, reachable or not
    n = 3 ** 9
    x = np.zeros(n, np.uint16)
    o = np.zeros(n, np.uint16)
    mask = np.zeros(n, np.uint16)
    over = np.zeros(n, np.uint8)
    winner = np.zeros(n, np.int8)
    for i in range(n):
        d, k, g = i, 0, np.zeros(9, int)
        while k < 9:
            g[k] = (0, 1, -1)[d % 3]
            d //= 3
            k += 1
        b = TB(g.reshape(3, 3))
        x[i] = sum(1 << k for k in range(9) if g[k] == 1)
        o[i] = sum(1 << k for k in range(9) if g[k] == -1)
        mask[i] = sum(1 << (r * 3 + c) for r, c in b.generate_possible_moves())
        ov, w = b.is_game_over()
        over[i] = ov
        winner[i] = 2 if w is None else w
        # make_move agrees with the mask: legal iff empty
        for k in (i % 9, (i * 7 + 3) % 9):
            try:
                nb = b.make_move(k // 3, k % 3, 1)
                assert g[k] == 0 and nb.board[k // 3, k % 3] == 1
            except ValueError:
                assert g[k] != 0
    # the reference's only golden data: tic_tac_toe_data.csv (canonical state, one-hot action)
    import csv

    st, ac = [], []
    with open(os.path.join(ref_shim.REF_ROOT, "tic_tac_toe_data.csv")) as f:
        rd = csv.reader(f)
        next(rd)
        for row in rd:
            st.append([int(v) for v in row[0].split()])
            ac.append([int(v) for v in row[1].split()])
    np.savez_compressed(os.path.join(OUT, "ttt_env.npz"), x=x, o=o, mask=mask, over=over, winner=winner,
                        csv_state=np.array(st, np.int8), csv_action=np.array(ac, np.int8))
    print(f"ttt: {n} boards, csv rows {len(st)}")


def mcts_fixtures(RB, TB):
    recs = {}
    C_PUCT = 1.25

    def record(prefix, game, roots, n_sims, salt):
        A = game.n_actions
        me, opp, cnt, W, P = [], [], [], [], []
        for board, player in roots:
            m = mr.MCTS(game, C_PUCT, lambda a, b: mr.hash_eval(a, b, salt, A))
            m.reset(board, player)
            m.run(n_sims)
            c, w, p = m.root_stats()
            a, b = game.wire(board, player)
            me.append(a)
            opp.append(b)
            cnt.append(c)
            W.append(w)
            P.append(p)
        recs[prefix + "_me"] = np.array(me, np.uint64)
        recs[prefix + "_opp"] = np.array(opp, np.uint64)
        recs[prefix + "_counts"] = np.stack(cnt)
        recs[prefix + "_W"] = np.stack(W)
        recs[prefix + "_P"] = np.stack(P)
        recs[prefix + "_meta"] = np.array([n_sims, salt], np.int64)
        print(prefix, len(roots), "roots", n_sims, "sims")

    # config 1: tic-tac-toe self-play with the reference loop order, 25 and 100 sims/move
    tg = mr.TicTacToeGame(TB)
    for n_sims in (25, 100):
        for salt in range(4):
            hist, winner = mr.self_play_game(tg, n_sims, C_PUCT, lambda a, b: mr.hash_eval(a, b, salt, 9))
            p = f"ttt_game_s{n_sims}_k{salt}"
            recs[p + "_me"] = np.array([h[0] for h in hist], np.uint64)
            recs[p + "_opp"] = np.array([h[1] for h in hist], np.uint64)
            recs[p + "_player"] = np.array([h[2] for h in hist], np.int8)
            recs[p + "_counts"] = np.stack([h[3] for h in hist])
            recs[p + "_action"] = np.array([h[4] for h in hist], np.uint8)
            recs[p + "_winner"] = np.array(winner)
    # Reversi 8x8: reachable roots along a random playout (includes late-game / pass / near-terminal)
    rng = np.random.default_rng(7)
    rg = mr.ReversiGame(RB, 8)
    roots = []
    b, cur, over = RB(size=8), 1, False
    while not over:
        roots.append((b, cur))
        moves = b.generate_possible_moves(cur)
        if moves:
            r, c = moves[int(rng.integers(len(moves)))]
            b = b.make_move(r, c, cur)
        over = b.is_game_over()
        cur *= -1
    record("rev8_playout_s48", rg, roots, 48, 3)
    record("rev8_start_s400", rg, [rg.initial()], 400, 11)
    # hunt for roots where the mover must pass
    pass_roots = []
    seed = 0
    while len(pass_roots) < 6 and seed < 400:
        rr = np.random.default_rng(1000 + seed)
        seed += 1
        b, cur, over = RB(size=8), 1, False
        while not over:
            moves = b.generate_possible_moves(cur)
            if not moves:
                pass_roots.append((b, cur))
            else:
                r, c = moves[int(rr.integers(len(moves)))]
                b = b.make_move(r, c, cur)
            over = b.is_game_over()
            cur *= -1
    record("rev8_pass_s64", rg, pass_roots[:6], 64, 5)
    # small boards: full deterministic self-play games
    for size, n_sims in ((4, 40), (6, 24)):
        g = mr.ReversiGame(RB, size)
        hist, winner = mr.self_play_game(g, n_sims, C_PUCT, lambda a, b: mr.hash_eval(a, b, 2, 65))
        p = f"rev{size}_game_s{n_sims}"
        recs[p + "_me"] = np.array([h[0] for h in hist], np.uint64)
        recs[p + "_opp"] = np.array([h[1] for h in hist], np.uint64)
        recs[p + "_player"] = np.array([h[2] for h in hist], np.int8)
        recs[p + "_counts"] = np.stack([h[3] for h in hist])
        recs[p + "_action"] = np.array([h[4] for h in hist], np.uint8)
        recs[p + "_winner"] = np.array(winner)
        print(p, len(hist), "plies, winner", winner)
    recs["c_puct"] = np.array(C_PUCT, np.float32)
    np.savez_compressed(os.path.join(OUT, "mcts.npz"), **recs)


def mcts_vl_fixtures(RB, TB):
    """Virtual-loss searches (MCTS.run_vl, K descents per iteration) of the definition on the LIVE reference boards:
    tests/golden/mcts_vl.npz.  Same root sets as mcts.npz's rev8_playout / start position, plus tic-tac-toe games."""
    recs = {}
    C_PUCT = 1.25

    def record(prefix, game, roots, n_sims, salt, leaves):
        A = game.n_actions
        me, opp, cnt, W, P = [], [], [], [], []
        for board, player in roots:
            m = mr.MCTS(game, C_PUCT, lambda a, b: mr.hash_eval(a, b, salt, A))
            m.reset(board, player)
            m.run_vl(n_sims, leaves)
            c, w, p = m.root_stats()
            a, b = game.wire(board, player)
            me.append(a)
            opp.append(b)
            cnt.append(c)
            W.append(w)
            P.append(p)
        recs[prefix + "_me"] = np.array(me, np.uint64)
        recs[prefix + "_opp"] = np.array(opp, np.uint64)
        recs[prefix + "_counts"] = np.stack(cnt)
        recs[prefix + "_W"] = np.stack(W)
        recs[prefix + "_P"] = np.stack(P)
        recs[prefix + "_meta"] = np.array([n_sims, salt, leaves], np.int64)
        print(prefix, len(roots), "roots", n_sims, "sims", leaves, "leaves")

    rng = np.random.default_rng(7)
    rg = mr.ReversiGame(RB, 8)
    roots = []
    b, cur, over = RB(size=8), 1, False
    while not over:
        roots.append((b, cur))
        moves = b.generate_possible_moves(cur)
        if moves:
            r, c = moves[int(rng.integers(len(moves)))]
            b = b.make_move(r, c, cur)
        over = b.is_game_over()
        cur *= -1
    record("rev8_playout_s48_k2", rg, roots, 48, 3, 2)
    record("rev8_playout_s48_k4", rg, roots[::2], 48, 3, 4)
    record("rev8_start_s240_k4", rg, [rg.initial()], 240, 11, 4)
    tg = mr.TicTacToeGame(TB)
    for leaves in (2, 4):
        hist, winner = mr.self_play_game(tg, 48, C_PUCT, lambda a, b: mr.hash_eval(a, b, 1, 9), leaves=leaves)
        p = f"ttt_game_s48_k{leaves}"
        recs[p + "_me"] = np.array([h[0] for h in hist], np.uint64)
        recs[p + "_opp"] = np.array([h[1] for h in hist], np.uint64)
        recs[p + "_counts"] = np.stack([h[3] for h in hist])
        recs[p + "_action"] = np.array([h[4] for h in hist], np.uint8)
        recs[p + "_winner"] = np.array(winner)
    recs["c_puct"] = np.array(C_PUCT, np.float32)
    np.savez_compressed(os.path.join(OUT, "mcts_vl.npz"), **recs)


def main():
    if not ref_shim.available():
        raise SystemExit("reference not found at " + ref_shim.REF_ROOT)
    os.makedirs(OUT, exist_ok=True)
    RB, TB = ref_shim.reversi_board_cls(), ref_shim.ttt_board_cls()
    if "--only-vl" in sys.argv:  # the other fixtures are unchanged
        mcts_vl_fixtures(RB, TB)
        return
    reversi_env(RB)
    ttt_env(TB)
    mcts_fixtures(RB, TB)
    mcts_vl_fixtures(RB, TB)


if __name__ == "__main__":
    main()
