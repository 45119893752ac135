"""Tree reuse across moves (opt-in: bz_mcts_reroot / BatchedMCTS.advance / BatchedSelfPlay(reuse=True)).

The definition is oracle/mcts_ref.py MCTS.advance (restated in oracle.c, orc_mcts_advance): after a move the search
continues on the subtree of the child the move leads to, iff that child has been expanded and the subtree fits the cap;
the next search adds its simulations to the kept statistics.  Checked here: the visit counts of every move of multi-move
play against the oracle, bit for bit (one leaf and four virtual-loss leaves per iteration, with and without a cap that
forces some trees to be dropped), the one-launch search against the per-iteration kernels on kept trees (node block by
node block), and lockstep self-play with reuse through both."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _early_roots(B, seed, size=8, max_plies=14):
    """reachable positions of the first plies of seeded random playouts (no game ends within the tests' few moves)"""
    from betazero_b200 import env
    from oracle import pyoracle as po

    rng = np.random.default_rng(seed)
    me, opp = np.empty(B, np.uint64), np.empty(B, np.uint64)
    for i in range(B):
        g, pl, _, _ = po.random_playout(size, int(seed * 7919 + i), int(rng.integers(0, max_plies)))
        me[i], opp[i] = po.grid_to_wire(g, pl)
    return me, opp, env.to_device_u64(me), env.to_device_u64(opp)


@pytest.mark.parametrize("leaves", [1, 4])
@pytest.mark.parametrize("size,cap", [(8, None), (8, 60), (6, None)])
def test_multi_move_play_with_reuse_matches_the_oracle(leaves, size, cap):
    from betazero_b200 import env, mcts
    from oracle import pyoracle as po

    B, n_sims, plies, salt = 40, 48, 6, 11
    me_h, opp_h, me, opp = _early_roots(B, seed=3 + size, size=size)
    cnt_o, act_o, kept_o = po.play_hash(me_h, opp_h, n_sims, plies, size=size, salt=salt, leaves=leaves, reuse=True,
                                        cap_units=-1 if cap is None else cap)
    assert (act_o != 255).all()  # no game ends inside the test
    pools = mcts.TreePools(B, n_sims, board_size=size, n_leaves=leaves, reuse=True)
    s = mcts.BatchedMCTS(pools, mcts.HashEvaluator(salt), use_graph=False)
    s.reset(me, opp)
    kept_seen = 0
    for p in range(plies):
        s.run(n_sims)
        s.check_errors()
        counts = s.root_policy()[0].cpu().numpy()
        assert np.array_equal(counts, cnt_o[:, p]), (p, leaves)
        a = s.best_action().clone()
        assert np.array_equal(a.cpu().numpy(), act_o[:, p])
        me, opp, err = env.apply(me, opp, a, size)
        assert not err.any()
        s.advance(a, me, opp, cap_units=cap)
        kept = (pools.inherited > 0).cpu().numpy()
        assert np.array_equal(kept, kept_o[:, p].astype(bool)), p
        # a kept root starts with the visits of the edge that led to it
        chosen = counts[np.arange(B), act_o[:, p]]
        assert np.array_equal(pools.inherited.cpu().numpy(), np.where(kept, chosen, 0))
        kept_seen += int(kept.sum())
    assert kept_seen > 0
    if cap is not None:
        assert kept_seen < B * plies  # the cap dropped some subtrees


def test_a_tree_is_emptied_when_the_new_position_is_not_the_childs():
    from betazero_b200 import env, mcts

    B = 8
    me, opp, _ = env.reversi_init(B)
    pools = mcts.TreePools(B, 64, n_leaves=4, reuse=True)
    s = mcts.BatchedMCTS(pools, mcts.HashEvaluator(1), use_graph=False)
    s.reset(me, opp)
    s.run(64)
    a = s.best_action().clone()
    me2, opp2, _ = env.apply(me, opp, a)
    other_me, other_opp = me2.clone(), opp2.clone()
    other_me[::2], other_opp[::2] = me[::2], opp[::2]  # even trees: "a new game started" (the start position again)
    s.advance(a, other_me, other_opp)
    inh = pools.inherited.cpu().numpy()
    assert (inh[::2] == 0).all() and (inh[1::2] > 0).all()
    assert (pools.arena_used.cpu().numpy()[::2] == 0).all()
    assert torch.equal(pools.root_me, other_me) and torch.equal(pools.root_opp, other_opp)
    # an action the root has no edge for also empties the tree
    s.run(64)
    bad = torch.full((B,), 0, dtype=torch.uint8, device="cuda")  # cell 0 is not a legal move in these positions
    s.advance(bad, me, opp)
    assert (pools.inherited == 0).all()
    with pytest.raises(RuntimeError):
        mcts.BatchedMCTS(mcts.TreePools(B, 8), mcts.HashEvaluator(1), use_graph=False).advance(a, me, opp)


def test_one_launch_search_on_kept_trees_equals_the_per_iteration_kernels():
    from betazero_b200 import env, mcts, net

    model = net.make_net("mlp", seed=3)
    B, n_sims, plies = 300, 64, 5
    _, _, me0, opp0 = _early_roots(B, seed=21)
    runs = []
    for one in (True, False):
        pools = mcts.TreePools(B, n_sims, n_leaves=4, reuse=True)
        pools.arena.zero_()
        pools.scratch.zero_()
        s = mcts.BatchedMCTS(pools, mcts.FusedNetEvaluator(model), use_graph=False, one_launch=one)
        assert s.one_launch is one
        me, opp = me0, opp0
        s.reset(me, opp)
        per_ply = []
        for _ in range(plies):
            s.run(n_sims)
            s.check_errors()
            per_ply.append([x.clone() for x in s.root_edges()] + [pools.arena_used.clone(), pools.sim_count.clone(),
                                                                  pools.depth_sum.clone(), pools.root_meta.clone()])
            a = s.best_action().clone()
            me, opp, err = env.apply(me, opp, a)
            assert not err.any()
            s.advance(a, me, opp)
        runs.append((per_ply, pools))
    for x, y in zip(runs[0][0], runs[1][0]):
        for u, v in zip(x, y):
            assert torch.equal(u, v)
    pa, pb = runs[0][1], runs[1][1]
    assert int(pa.inherited.sum()) > 0 and torch.equal(pa.inherited, pb.inherited)
    used = pa.arena_used.long() * 8
    arena_a, arena_b = pa.arena.view(B, -1), pb.arena.view(B, -1)
    live = torch.arange(arena_a.shape[1], device="cuda")[None, :] < used[:, None]
    assert torch.equal(arena_a[live], arena_b[live])  # every node block of every kept tree


def test_selfplay_with_reuse_plays_the_same_games_through_both_search_paths_and_digs_deeper():
    from betazero_b200 import mcts, net, selfplay

    model = net.make_net("mlp", seed=2)
    recs, depth = [], {}
    for reuse, one in ((True, True), (True, False), (False, True)):
        sp = selfplay.BatchedSelfPlay(96, 32, mcts.FusedNetEvaluator(model), board_size=6, temp_plies=6, seed=11, n_leaves=4,
                                      use_graph=False, one_launch=one, reuse=reuse)
        roots = []
        for _ in range(44):
            sp.play_move()
            roots.append(int(sp.pools.sim_count.sum()))
        sp.mcts.check_errors()
        st = sp.stats()
        assert st["games"] >= 96 and st["dropped"] == 0
        r = sp.drain_replay()
        order = torch.argsort(r["game"] * 256 + r["ply"].long())
        if reuse:
            recs.append((st, {k: v[order].cpu() for k, v in r.items()}))
        depth[reuse] = max(roots)
    assert recs[0][0] == recs[1][0]
    for k in recs[0][1]:
        assert torch.equal(recs[0][1][k], recs[1][1][k]), k
    assert depth[True] > depth[False]  # kept trees: the roots carry more than one search's visits


@pytest.mark.parametrize("leaves", [1, 4])
def test_reference_style_players_with_reuse_follow_the_definition(leaves):
    """Two MCTSPlayer(reuse=True) in the reference's episode loop (reversi_terminal.py:16-38): each keeps its own tree
    across its moves, following its move and the opponent's reply (or pass); the visit counts of every get_move equal
    those of the sequential definition advanced the same way (MCTS.advance twice per own turn)."""
    from betazero_b200 import boards, players
    from oracle import mcts_ref as mr
    from oracle import pyoracle as po

    size, n_sims, salt = 6, 32, 4
    game = mr.ReversiGame(po.OracleReversiBoard, size)
    ev = lambda a, b: mr.hash_eval(a, b, salt, 65)
    gpu = {s: players.MCTSPlayer(s, n_sims=n_sims, size=size, salt=salt, n_leaves=leaves, reuse=True) for s in (1, -1)}
    orc = {s: None for s in (1, -1)}
    board = boards.ReversiBoard(size=size)
    ob, player = game.initial()
    kept_moves = 0
    for ply in range(26):
        if ob.is_game_over():
            break
        moves = ob.generate_possible_moves(player)
        if moves:
            if orc[player] is None:
                orc[player] = mr.MCTS(game, 1.25, ev)
                orc[player].reset(ob, player)
            m = orc[player]
            inherited = int(m.root.visits if leaves > 1 else (1 + int(np.sum(m.root.N))) if m.root.actions is not None else 0)
            (m.run_vl(n_sims, leaves) if leaves > 1 else m.run(n_sims))
            cnt = m.root_stats()[0]
            row, col = gpu[player].get_move(board)
            assert np.array_equal(gpu[player].last_counts, cnt), ply
            assert gpu[player].last_inherited == inherited
            kept_moves += inherited > 0
            a = mr.pick_move(cnt)
            assert (row, col) == (a >> 3, a & 7)
            board = board.make_move(row, col, player)
        else:
            a = 64  # the loop skips a player without a move: a pass nobody's get_move sees
            orc[player] = None  # ... and a player that was skipped sees two new discs at its next turn: it starts over
        ob, nxt = game.next(ob, player, a)
        for s in (1, -1):  # both players' trees follow the move
            if orc[s] is not None:
                orc[s].advance(a, ob, nxt, gpu[s].pools.reuse_cap_units)
        player = nxt
    assert kept_moves >= 4


@pytest.mark.parametrize("leaves", [1, 2])
def test_tree_reuse_on_tic_tac_toe_matches_the_oracle(leaves):
    """bz_mcts_reroot is game-agnostic (it moves node blocks): config-1 style tic-tac-toe play from the empty board and
    from a few two-ply positions, every move's visit counts against orc_mcts_play_hash with reuse"""
    from betazero_b200 import mcts
    from oracle import pyoracle as po

    starts = [(0, 0), (1 << 4, 1 << 0), (1 << 0, 1 << 8), (1 << 2, 1 << 4), (1 << 7, 1 << 1)]  # (me, opp): X to move
    me_h = np.array([s[0] for s in starts], np.uint64)
    opp_h = np.array([s[1] for s in starts], np.uint64)
    B, n_sims, plies, salt = len(starts), 40, 5, 3
    cnt_o, act_o, kept_o = po.play_hash(me_h, opp_h, n_sims, plies, game=po.GAME_TTT, salt=salt, leaves=leaves, reuse=True)
    pools = mcts.TreePools(B, n_sims, game=mcts.GAME_TTT, n_leaves=leaves, reuse=True)
    s = mcts.BatchedMCTS(pools, mcts.HashEvaluator(salt), use_graph=False)
    me = torch.from_numpy(me_h.view(np.int64)).cuda()
    opp = torch.from_numpy(opp_h.view(np.int64)).cuda()
    s.reset(me, opp)
    alive = np.ones(B, bool)
    for p in range(plies):
        s.run(n_sims)
        s.check_errors()
        counts = s.root_policy()[0].cpu().numpy()
        alive &= act_o[:, p] != 255  # games that have ended: the oracle stops, the engine searches a terminal root
        assert np.array_equal(counts[alive], cnt_o[alive, p]), p
        a = s.best_action().clone()
        assert np.array_equal(a.cpu().numpy()[alive], act_o[alive, p])
        bit = torch.where(a < 9, torch.ones_like(me) << a.long().clamp(max=8), torch.zeros_like(me))
        me, opp = opp.clone(), (me | bit)  # the mover's mark is placed, the view flips to the next mover
        s.advance(a, me, opp)
        kept = (pools.inherited > 0).cpu().numpy()
        assert np.array_equal(kept[alive], kept_o[alive, p].astype(bool)), p
    assert kept_o.sum() > 0
