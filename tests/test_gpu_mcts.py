"""GPU parity: batched MCTS kernels (K5-K8) through the C ABI vs the sequential oracle.

Visit counts must be BIT-EXACT (deterministic tie-breaking, Dirichlet noise off); W / P are
compared exactly too (they are, by construction of the fp32 op order), and Q / pi within the
north-star's 1e-5 relative tolerance.  MCTS parity is unpinned by the reference (it has no MCTS):
the oracle is oracle/mcts_ref.py (golden) and its C restatement (oracle.c)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star: Q values and policy targets within 1e-5 relative


GROUPS = [32, 16, 8]  # lanes per tree: warp-per-tree (small batches), 2 and 4 trees per warp (larger batches)


def _search(me_h, opp_h, n_sims, salt, game=0, size=8, c_puct=1.25, group_lanes=0, **kw):
    from betazero_b200 import env, mcts

    pools = mcts.TreePools(len(me_h), n_sims, game=game, board_size=size, c_puct=c_puct, group_lanes=group_lanes)
    s = mcts.BatchedMCTS(pools, mcts.HashEvaluator(salt), **kw)
    cnt, pi, q = s.search(env.to_device_u64(me_h), env.to_device_u64(opp_h), n_sims)
    return s, cnt.cpu().numpy(), pi.cpu().numpy(), q.cpu().numpy()


def _root_W_P(s):
    """root-edge W and P, scattered by action"""
    _, W, P = s.root_edges()
    return W.cpu().numpy(), P.cpu().numpy()


@pytest.mark.parametrize("group", GROUPS)
@pytest.mark.parametrize("prefix", ["rev8_playout_s48", "rev8_start_s400", "rev8_pass_s64"])
def test_reversi_matches_golden(golden_mcts, prefix, group):
    g = golden_mcts
    n_sims, salt = (int(v) for v in g[prefix + "_meta"])
    s, cnt, pi, q = _search(g[prefix + "_me"], g[prefix + "_opp"], n_sims, salt, c_puct=float(g["c_puct"]),
                            group_lanes=group)
    assert np.array_equal(cnt, g[prefix + "_counts"])
    W, P = _root_W_P(s)
    assert np.array_equal(W, g[prefix + "_W"]) and np.array_equal(P, g[prefix + "_P"])
    gc = g[prefix + "_counts"].astype(np.float32)
    exp_pi = gc / gc.sum(1, keepdims=True)
    exp_q = np.where(gc > 0, g[prefix + "_W"] / np.maximum(gc, 1), 0)
    np.testing.assert_allclose(pi, exp_pi, rtol=RTOL, atol=0)
    np.testing.assert_allclose(q, exp_q, rtol=RTOL, atol=0)


@pytest.mark.parametrize("group", GROUPS)
@pytest.mark.parametrize("n_sims", [25, 100])
def test_config1_ttt_selfplay_visit_counts(golden_mcts, n_sims, group):
    """BASELINE config 1: tic-tac-toe MCTS self-play, root visit counts of every ply bit-exact"""
    from betazero_b200 import mcts

    g = golden_mcts
    for salt in range(4):
        p = f"ttt_game_s{n_sims}_k{salt}"
        s, cnt, pi, q = _search(g[p + "_me"], g[p + "_opp"], n_sims, salt, game=mcts.GAME_TTT, c_puct=float(g["c_puct"]),
                                group_lanes=group)
        assert np.array_equal(cnt, g[p + "_counts"])
        assert np.array_equal(s.best_action().cpu().numpy(), g[p + "_action"])


@pytest.mark.parametrize("group", GROUPS)
@pytest.mark.parametrize("size,n_sims", [(4, 40), (6, 24)])
def test_small_boards_match_golden(golden_mcts, size, n_sims, group):
    g = golden_mcts
    p = f"rev{size}_game_s{n_sims}"
    s, cnt, _, _ = _search(g[p + "_me"], g[p + "_opp"], n_sims, 2, size=size, c_puct=float(g["c_puct"]), group_lanes=group)
    assert np.array_equal(cnt, g[p + "_counts"])
    assert np.array_equal(s.best_action().cpu().numpy(), g[p + "_action"])


@pytest.mark.parametrize("group", GROUPS)
def test_config3_1024_trees_100_sims_vs_c_oracle(group):
    """BASELINE config 3 shape: 1024 concurrent trees, 100 sims/move (hash evaluator for parity)."""
    from oracle import pyoracle as po

    me_h, opp_h = po.playout_boards(1024, seed=5)
    s, cnt, pi, q = _search(me_h, opp_h, 100, salt=17, group_lanes=group)
    r_cnt, r_W, r_P, ctr = po.search_hash(me_h, opp_h, 100, po.GAME_REVERSI, 8, 1.25, 17)
    assert np.array_equal(cnt, r_cnt)
    W, P = _root_W_P(s)
    assert np.array_equal(W, r_W) and np.array_equal(P, r_P)
    st = s.stats()
    assert st["sims"] == ctr["sims"] == 1024 * 100
    assert abs(st["mean_depth"] - ctr["sum_depth"] / ctr["sims"]) < 1e-9
    assert st["edges"] == ctr["edges"]


@pytest.mark.parametrize("group", GROUPS)
def test_800_sims_vs_c_oracle(group):
    """the headline search length (800 sims/move) on 128 trees, graph + fused path"""
    from oracle import pyoracle as po

    me_h, opp_h = po.playout_boards(128, seed=6)
    s, cnt, _, _ = _search(me_h, opp_h, 800, salt=1, group_lanes=group)
    r_cnt, _, _, _ = po.search_hash(me_h, opp_h, 800, po.GAME_REVERSI, 8, 1.25, 1)
    assert np.array_equal(cnt, r_cnt)


def test_headline_size_4096_trees_800_sims():
    """BASELINE configs[3] size: 4096 trees x 800 iterations.  The C oracle checks a 160-tree sample
    bit for bit (trees are independent); invariants are checked on every tree."""
    from oracle import pyoracle as po

    B, S = 4096, 800
    me_h, opp_h = po.playout_boards(B, seed=44)
    s, cnt, pi, q = _search(me_h, opp_h, S, salt=9)
    live = po.terminal(me_h, opp_h)[0] == 0
    assert (cnt.sum(1)[live] == S - 1).all() and (cnt.sum(1)[~live] == 0).all()
    np.testing.assert_allclose(pi.sum(1)[live], 1.0, atol=1e-5)
    mask = po.legal_mask(me_h, opp_h)
    legal = ((mask[:, None] >> np.arange(64, dtype=np.uint64)) & np.uint64(1)).astype(bool)
    assert (cnt[:, :64][~legal] == 0).all()  # no visits on illegal cells
    assert ((cnt[:, 64] > 0) == ((mask == 0) & live)).all()  # pass is searched exactly when forced
    assert (np.abs(q) <= 1.0).all()
    st = s.stats()
    assert st["sims"] == int(live.sum()) * S + int((~live).sum()) * S  # finished roots still count iterations
    sample = np.arange(0, B, B // 160)
    r_cnt, r_W, r_P, _ = po.search_hash(me_h[sample], opp_h[sample], S, po.GAME_REVERSI, 8, 1.25, 9)
    assert np.array_equal(cnt[sample], r_cnt)
    W, P = _root_W_P(s)
    assert np.array_equal(W[sample], r_W) and np.array_equal(P[sample], r_P)


def test_headline_config_k4_wave_4096_trees_800_sims():
    """The EXACT configuration bench.py times (BASELINE configs[3]): 4096 trees x 800 sims/move with 4 virtual-loss
    descents per tree and iteration (step_wave_kernel<reversi, 8>, CUDA graph + fused step).  EVERY tree is checked
    bit for bit (N, W, P of the root) against the C oracle's select_vl definition."""
    from oracle import pyoracle as po

    B, S, K = 4096, 800, 4
    me_h, opp_h = po.playout_boards(B, seed=45)
    s, cnt, pi, q = _search_vl(me_h, opp_h, S, 11, K)
    live = po.terminal(me_h, opp_h)[0] == 0
    # the K descents of the first iteration all end on the unexpanded root: S - K visits on the root edges
    assert (cnt.sum(1)[live] == S - K).all() and (cnt.sum(1)[~live] == 0).all()
    np.testing.assert_allclose(pi.sum(1)[live], 1.0, atol=1e-5)
    mask = po.legal_mask(me_h, opp_h)
    legal = ((mask[:, None] >> np.arange(64, dtype=np.uint64)) & np.uint64(1)).astype(bool)
    assert (cnt[:, :64][~legal] == 0).all()
    assert ((cnt[:, 64] > 0) == ((mask == 0) & live)).all()
    assert (np.abs(q) <= 1.0).all()
    sample = np.arange(B)  # every tree: the C oracle needs ~2 s for 4096 x 800 simulations
    r_cnt, r_W, r_P, ctr = po.search_hash(me_h[sample], opp_h[sample], S, po.GAME_REVERSI, 8, 1.25, 11, leaves=K)
    assert np.array_equal(cnt[sample], r_cnt)
    W, P = _root_W_P(s)
    assert np.array_equal(W[sample], r_W) and np.array_equal(P[sample], r_P)
    np.testing.assert_allclose(q[sample], np.where(r_cnt > 0, r_W / np.maximum(r_cnt, 1), 0), rtol=RTOL, atol=0)


def test_large_batch_uses_eight_lane_groups():
    """>= 8192 trees: the 4-trees-per-warp kernels are selected automatically; same answers"""
    from oracle import pyoracle as po

    B, S = 8192 + 5, 60
    me_h, opp_h = po.playout_boards(512, seed=3)
    me_h, opp_h = np.resize(me_h, B), np.resize(opp_h, B)
    s, cnt, _, _ = _search(me_h, opp_h, S, salt=2)
    r_cnt, _, _, _ = po.search_hash(me_h[:512], opp_h[:512], S, po.GAME_REVERSI, 8, 1.25, 2)
    assert np.array_equal(cnt[:512], r_cnt)
    assert np.array_equal(cnt[512:1024], r_cnt) and np.array_equal(cnt[-5:], r_cnt[(B - 5) % 512:][:5])


def test_fused_graph_and_plain_paths_agree():
    from oracle import pyoracle as po

    me_h, opp_h = po.playout_boards(300, seed=8)
    ref = None
    for kw in (dict(use_graph=False, fused=False), dict(use_graph=False, fused=True),
               dict(use_graph=True, fused=True, graph_unroll=8), dict(use_graph=True, fused=False, graph_unroll=5)):
        _, cnt, pi, q = _search(me_h, opp_h, 70, salt=3, **kw)
        if ref is None:
            ref = (cnt, pi, q)
        else:
            assert np.array_equal(cnt, ref[0]) and np.array_equal(pi, ref[1]) and np.array_equal(q, ref[2])


@pytest.mark.parametrize("group", GROUPS)
def test_lockstep_float_priors_vs_c_oracle(group):
    """real-net-like float priors and values: feed the SAME (w, v) to the GPU trees and to the
    oracle trees at every iteration and compare leaves, statuses and final statistics bit for bit"""
    from betazero_b200 import env, mcts
    from oracle import pyoracle as po

    B, n_sims = 96, 150
    me_h, opp_h = po.playout_boards(B, seed=12)
    pools = mcts.TreePools(B, n_sims, c_puct=2.0, group_lanes=group)
    s = mcts.BatchedMCTS(pools, None, use_graph=False)
    s.reset(env.to_device_u64(me_h), env.to_device_u64(opp_h))
    trees = [po.OracleTree(po.GAME_REVERSI, 8, 2.0) for _ in range(B)]
    for t, m, o in zip(trees, me_h, opp_h):
        t.reset_wire(m, o)
    rng = np.random.default_rng(0)
    s.select()
    for it in range(n_sims):
        lm, lo = env.to_host_u64(pools.leaf_me), env.to_host_u64(pools.leaf_opp)
        st = pools.leaf_status.cpu().numpy()
        plen = pools.path_len.cpu().numpy()
        w = (rng.random((B, 65)) ** 4).astype(np.float32)
        w[rng.random((B, 65)) < 0.1] = 0.0  # exact zeros happen with bf16 softmax
        v = rng.uniform(-1, 1, B).astype(np.float32)
        for i, t in enumerate(trees):
            ost, ome, oopp, od = t.select()
            assert (ost, ome, oopp, od) == (int(st[i]), int(lm[i]), int(lo[i]), int(plen[i])), (it, i)
            t.expand_backup(w[i], v[i])
        s.prior_w.copy_(torch.from_numpy(w))
        s.value.copy_(torch.from_numpy(v))
        if it + 1 < n_sims:
            s.step()
        else:
            s.expand_backup()
    cnt, pi, q = (x.cpu().numpy() for x in s.root_policy())
    W, P = _root_W_P(s)
    for i, t in enumerate(trees):
        c, w_, p_ = t.root_stats()
        assert np.array_equal(cnt[i], c) and np.array_equal(W[i], w_) and np.array_equal(P[i], p_)
    s.check_errors()


def test_leaf_planes_are_canonical_leaf_boards():
    from betazero_b200 import env, mcts
    from oracle import pyoracle as po

    me_h, opp_h = po.playout_boards(257, seed=2)
    pools = mcts.TreePools(257, 40)
    s = mcts.BatchedMCTS(pools, mcts.HashEvaluator(4), use_graph=False)
    s.reset(env.to_device_u64(me_h), env.to_device_u64(opp_h))
    s.run(39)
    s.select()
    exp = env.planes(pools.leaf_me, pools.leaf_opp)
    assert torch.equal(pools.leaf_planes, exp)
    pools.leaf_planes.zero_()
    from betazero_b200 import _lib
    _lib.check(_lib.load().bz_mcts_gather(pools._ref, _lib.stream_ptr()))
    assert torch.equal(pools.leaf_planes, exp)


def test_terminal_and_pass_roots():
    from betazero_b200 import env, mcts

    # tree 0: full board (game over).  tree 1: mover owns b1, opponent a1, rest empty: the mover
    # cannot outflank a corner disc, but the opponent can play c1 -> the mover must pass.
    me_h = np.array([(1 << 64) - 1 - 0xFF, 0x2], np.uint64)
    opp_h = np.array([0xFF, 0x1], np.uint64)
    pools = mcts.TreePools(2, 16)
    s = mcts.BatchedMCTS(pools, mcts.HashEvaluator(0), use_graph=False)
    cnt, pi, q = s.search(env.to_device_u64(me_h), env.to_device_u64(opp_h), 16)
    cnt = cnt.cpu().numpy()
    assert cnt[0].sum() == 0  # finished game: nothing to search
    best = s.best_action().cpu().numpy()
    assert best[0] == 255
    assert cnt[1, 64] == 15 and cnt[1, :64].sum() == 0 and best[1] == 64
    assert pi.cpu().numpy()[1, 64] == 1.0


@pytest.mark.parametrize("leaves,group", [(1, 0), (4, 0), (2, 0), (3, 0), (2, 8)])
def test_arena_overflow_is_detected(leaves, group):
    """an arena too small for the search: flagged and raised in every kernel family (one leaf, wave mode, slots one
    after the other), and the trees stay readable (no write past the arena)"""
    from betazero_b200 import _lib, env, mcts
    from oracle import pyoracle as po

    me_h, opp_h = po.playout_boards(8, seed=1)
    pools = mcts.TreePools(8, 72, arena_units=40, n_leaves=leaves, group_lanes=group)
    guard = pools.arena.clone()  # the arena of the LAST tree must not be overrun into whatever follows: check its size
    s = mcts.BatchedMCTS(pools, mcts.HashEvaluator(0), use_graph=False)
    with pytest.raises(_lib.BzError, match="overflow"):
        s.search(env.to_device_u64(me_h), env.to_device_u64(opp_h), 72)
    assert pools.arena.numel() == guard.numel()
    assert int(pools.arena_used.max().item()) <= 40
    cnt, _, _ = s.root_policy()
    assert int(cnt.sum(1).max().item()) <= 72


def test_hash_eval_kernel_matches_oracle():
    from betazero_b200 import _lib, env
    from oracle import pyoracle as po

    me_h, opp_h = po.synthetic_boards(500, seed=21)
    me_d, opp_d = env.to_device_u64(me_h), env.to_device_u64(opp_h)  # keep alive: raw pointers are passed
    for A in (9, 65):
        w = torch.empty((500, A), dtype=torch.float32, device="cuda")
        v = torch.empty(500, dtype=torch.float32, device="cuda")
        _lib.check(_lib.load().bz_hash_eval(_lib.dptr(me_d), _lib.dptr(opp_d), 77, A,
                                            _lib.dptr(w), _lib.dptr(v), 500, _lib.stream_ptr()))
        w, v = w.cpu().numpy(), v.cpu().numpy()
        for i in range(0, 500, 7):
            rw, rv = po.hash_eval(me_h[i], opp_h[i], 77, A)
            assert np.array_equal(w[i], rw) and v[i] == rv


@pytest.mark.parametrize("group", GROUPS)
def test_logits_mode_fused_softmax_and_tanh(group):
    """BZ_PRIOR_LOGITS_BF16: the tree kernel does the legal-move softmax and tanh itself (ex2.approx for the
    softmax, libm tanhf for the value): priors and values within the north-star's 1e-5 RELATIVE of float64 numpy."""
    from betazero_b200 import env, mcts
    from oracle import pyoracle as po

    B = 512
    me_h, opp_h = po.playout_boards(B, seed=31)
    pools = mcts.TreePools(B, 4, prior_mode=mcts.PRIOR_LOGITS_BF16, eval_stride=72, group_lanes=group)

    class Raw:
        prior_mode = mcts.PRIOR_LOGITS_BF16
        stride = 72

        def bind(self, pools):
            g = torch.Generator(device="cuda").manual_seed(5)
            self.out = (torch.randn((B, 72), device="cuda", generator=g) * 3).to(torch.bfloat16)
            self.value = torch.zeros(1, device="cuda")
            return self.out, self.value

        def __call__(self, pools):
            pass

    ev = Raw()
    s = mcts.BatchedMCTS(pools, ev, use_graph=False)
    s.reset(env.to_device_u64(me_h), env.to_device_u64(opp_h))
    s.run(2)  # iteration 1 expands the root, iteration 2 visits one child and backs its value up
    N, W, P = (x.cpu().numpy() for x in s.root_edges())
    logits = ev.out.float().cpu().numpy()
    mask = po.legal_mask(me_h, opp_h)
    over = po.terminal(me_h, opp_h)[0]
    for i in range(B):
        if over[i]:  # finished game: nothing is expanded
            assert N[i].sum() == 0 and P[i].sum() == 0
            continue
        legal = [a for a in range(64) if (int(mask[i]) >> a) & 1] or [64]
        l = logits[i, legal].astype(np.float64)
        p = np.exp(l - l.max())
        p /= p.sum()
        np.testing.assert_allclose(P[i, legal], p, rtol=RTOL, atol=1e-12)  # 1e-5 RELATIVE (ex2.approx: 2^-22)
        assert P[i].sum() == pytest.approx(1.0, abs=1e-5) and N[i].sum() == 1
        a = int(np.argmax(N[i]))
        # iteration 2 visited child `a`.  A finished game backs up its exact score; any other child was evaluated from
        # the same logits row, and its value tanh(l[65]) came back negated: 1e-5 relative (north-star bound on Q)
        cm, co, cerr = po.apply(me_h[i:i + 1], opp_h[i:i + 1], np.array([a], np.uint8))
        assert not cerr[0]
        c_over, c_win, _, _ = po.terminal(cm, co)
        if c_over[0]:
            assert W[i, a] == -float(c_win[0])
        else:
            np.testing.assert_allclose(W[i, a], -np.tanh(np.float64(logits[i, 65])), rtol=RTOL, atol=0)


def test_fused_net_evaluator_agrees_with_parity_evaluator():
    """same net through the fast path (bf16 logits -> in-kernel softmax) and through the parity
    path (fp32 softmax weights): root priors agree to bf16 precision and the searches are close"""
    from betazero_b200 import env, mcts, net
    from oracle import pyoracle as po

    B, n_sims = 256, 64
    me_h, opp_h = po.playout_boards(B, seed=77)
    model = net.make_net("mlp", seed=3)
    res = []
    for ev in (mcts.NetEvaluator(model), mcts.FusedNetEvaluator(model)):
        pools = mcts.TreePools(B, n_sims)
        s = mcts.BatchedMCTS(pools, ev, use_graph=True, graph_unroll=8)
        cnt, pi, q = s.search(env.to_device_u64(me_h), env.to_device_u64(opp_h), n_sims)
        _, _, P = s.root_edges()
        res.append((cnt.cpu().numpy(), P.cpu().numpy()))
    np.testing.assert_allclose(res[0][1], res[1][1], atol=2e-2)
    live = po.terminal(me_h, opp_h)[0] == 0  # finished games have nothing to search
    assert (res[0][0].sum(1)[live] == n_sims - 1).all() and (res[1][0].sum(1)[live] == n_sims - 1).all()
    assert (res[0][0].sum(1)[~live] == 0).all()
    same = (res[0][0].argmax(1) == res[1][0].argmax(1))[live].mean()
    assert same > 0.8


def test_root_dirichlet_noise_formula_and_off_switch():
    """exploration noise (self-play only): P <- (1-eps) P + eps * noise/sum(noise over legal)"""
    from betazero_b200 import _lib, env, mcts
    from oracle import pyoracle as po

    B = 300
    me_h, opp_h = po.playout_boards(B, seed=14)
    pools = mcts.TreePools(B, 8)
    s = mcts.BatchedMCTS(pools, mcts.HashEvaluator(6), use_graph=False)
    s.reset(env.to_device_u64(me_h), env.to_device_u64(opp_h))
    s.run(1)  # expand the roots
    _, _, P0 = (x.cpu().numpy() for x in s.root_edges())
    g = torch.Generator(device="cuda").manual_seed(1)
    noise = torch.rand((B, 65), device="cuda", generator=g) + 0.01
    _lib.check(_lib.load().bz_mcts_root_noise(pools._ref, _lib.dptr(noise), 0.25, _lib.stream_ptr()))
    _, _, P1 = (x.cpu().numpy() for x in s.root_edges())
    nz = noise.cpu().numpy()
    legal = P0 > 0
    exp = np.where(legal, 0.75 * P0 + 0.25 * nz / np.maximum((nz * legal).sum(1, keepdims=True), 1e-30), 0)
    np.testing.assert_allclose(P1, exp, rtol=1e-5, atol=1e-7)
    live = legal.any(1)
    np.testing.assert_allclose(P1.sum(1)[live], 1.0, atol=1e-5)
    # through the search API: alpha > 0 changes visit counts, alpha == 0 is bit-identical to the oracle
    a = _search(me_h, opp_h, 64, salt=6)[1]
    pools2 = mcts.TreePools(B, 64)
    s2 = mcts.BatchedMCTS(pools2, mcts.HashEvaluator(6), use_graph=True, graph_unroll=8, dirichlet_alpha=0.3, noise_seed=5)
    b = s2.search(env.to_device_u64(me_h), env.to_device_u64(opp_h), 64)[0].cpu().numpy()
    assert (a.sum(1) == b.sum(1)).all() and (a != b).any()
    r_cnt, _, _, _ = po.search_hash(me_h, opp_h, 64, po.GAME_REVERSI, 8, 1.25, 6)
    assert np.array_equal(a, r_cnt)


def test_resnet_through_net_evaluator():
    from betazero_b200 import env, mcts, net
    from oracle import pyoracle as po

    B = 128
    me_h, opp_h = po.playout_boards(B, seed=4)
    model = net.make_net("resnet", channels=32, blocks=2, seed=0)
    pools = mcts.TreePools(B, 24)
    s = mcts.BatchedMCTS(pools, mcts.NetEvaluator(model), use_graph=True, graph_unroll=4)
    cnt, pi, q = s.search(env.to_device_u64(me_h), env.to_device_u64(opp_h), 24)
    live = po.terminal(me_h, opp_h)[0] == 0
    assert (cnt.cpu().numpy().sum(1)[live] == 23).all()
    assert torch.isfinite(pi).all() and torch.isfinite(q).all()


# ---- K leaves per iteration with virtual loss (pools.n_leaves > 1) --------------------------------------------
def _search_vl(me_h, opp_h, n_sims, salt, leaves, game=0, size=8, c_puct=1.25, group_lanes=0, **kw):
    from betazero_b200 import env, mcts

    pools = mcts.TreePools(len(me_h), n_sims, game=game, board_size=size, c_puct=c_puct, group_lanes=group_lanes,
                           n_leaves=leaves)
    s = mcts.BatchedMCTS(pools, mcts.HashEvaluator(salt), **kw)
    cnt, pi, q = s.search(env.to_device_u64(me_h), env.to_device_u64(opp_h), n_sims)
    return s, cnt.cpu().numpy(), pi.cpu().numpy(), q.cpu().numpy()


@pytest.mark.parametrize("group", GROUPS)
@pytest.mark.parametrize("leaves", [2, 4])
def test_virtual_loss_reversi_vs_c_oracle(group, leaves):
    """K descents per iteration with virtual loss (oracle.c Part 2d / mcts_ref.MCTS.select_vl): visit counts, W and P
    of the root bit for bit, and the search counters (the collisions are part of the definition)"""
    from oracle import pyoracle as po

    me_h, opp_h = po.playout_boards(512, seed=15)
    n_sims = 96
    s, cnt, pi, q = _search_vl(me_h, opp_h, n_sims, 5, leaves, group_lanes=group)
    r_cnt, r_W, r_P, ctr = po.search_hash(me_h, opp_h, n_sims, po.GAME_REVERSI, 8, 1.25, 5, leaves=leaves)
    assert np.array_equal(cnt, r_cnt)
    W, P = _root_W_P(s)
    assert np.array_equal(W, r_W) and np.array_equal(P, r_P)
    st = s.stats()
    assert st["sims"] == ctr["sims"] == 512 * n_sims
    assert st["edges"] == ctr["edges"]
    assert abs(st["mean_depth"] - ctr["sum_depth"] / ctr["sims"]) < 1e-9
    live = po.terminal(me_h, opp_h)[0] == 0
    assert (cnt.sum(1)[live] == n_sims - leaves).all()  # the first iteration's descents all end on the unexpanded root


@pytest.mark.parametrize("group", GROUPS)
def test_virtual_loss_ttt_and_small_boards(group):
    from betazero_b200 import mcts
    from oracle import pyoracle as po

    x = np.array([0, 0b000010000, 0b100000001, 0b000000111], dtype=np.uint64)
    o = np.array([0, 0b000000001, 0b000010010, 0b011000000], dtype=np.uint64)
    s, cnt, _, _ = _search_vl(x, o, 60, 3, 2, game=mcts.GAME_TTT, group_lanes=group)
    r_cnt, r_W, _, _ = po.search_hash(x, o, 60, po.GAME_TTT, 3, 1.25, 3, leaves=2)
    assert np.array_equal(cnt, r_cnt)
    assert np.array_equal(_root_W_P(s)[0], r_W)
    for size in (4, 6):
        me_h, opp_h = po.playout_boards(64, seed=size, size=size)
        s, cnt, _, _ = _search_vl(me_h, opp_h, 48, 1, 2, size=size, group_lanes=group)
        r_cnt, _, _, _ = po.search_hash(me_h, opp_h, 48, po.GAME_REVERSI, size, 1.25, 1, leaves=2)
        assert np.array_equal(cnt, r_cnt)


def test_virtual_loss_800_sims_graph_path_vs_c_oracle():
    """headline search length with 2 leaves per iteration (400 iterations), CUDA graph + fused step kernel"""
    from oracle import pyoracle as po

    me_h, opp_h = po.playout_boards(128, seed=6)
    s, cnt, _, _ = _search_vl(me_h, opp_h, 800, 1, 2)
    r_cnt, r_W, _, _ = po.search_hash(me_h, opp_h, 800, po.GAME_REVERSI, 8, 1.25, 1, leaves=2)
    assert np.array_equal(cnt, r_cnt)
    assert np.array_equal(_root_W_P(s)[0], r_W)
    # plain (unfused, no graph) launches give the same trees
    s2, cnt2, _, _ = _search_vl(me_h, opp_h, 800, 1, 2, use_graph=False, fused=False)
    assert np.array_equal(cnt2, cnt)


@pytest.mark.parametrize("leaves", [2, 3])
def test_virtual_loss_lockstep_float_priors_vs_c_oracle(leaves):
    """the same float (w, v) fed to GPU trees and oracle trees every iteration: leaves, statuses, path lengths of every
    slot and the final statistics must agree bit for bit"""
    from betazero_b200 import env, mcts
    from oracle import pyoracle as po

    B, iters = 64, 60
    me_h, opp_h = po.playout_boards(B, seed=12)
    pools = mcts.TreePools(B, iters * leaves, c_puct=2.0, n_leaves=leaves)
    s = mcts.BatchedMCTS(pools, None, use_graph=False)
    s.reset(env.to_device_u64(me_h), env.to_device_u64(opp_h))
    trees = [po.OracleTree(po.GAME_REVERSI, 8, 2.0) for _ in range(B)]
    for t, m, o in zip(trees, me_h, opp_h):
        t.reset_wire(m, o)
    rng = np.random.default_rng(1)
    s.select()
    for it in range(iters):
        lm, lo = env.to_host_u64(pools.leaf_me), env.to_host_u64(pools.leaf_opp)
        st = pools.leaf_status.cpu().numpy()
        plen = pools.path_len.cpu().numpy()
        w = (rng.random((leaves * B, 65)) ** 4).astype(np.float32)
        w[rng.random((leaves * B, 65)) < 0.1] = 0.0
        v = rng.uniform(-1, 1, leaves * B).astype(np.float32)
        for i, t in enumerate(trees):
            for j in range(leaves):
                r = j * B + i
                ost, ome, oopp, od = t.select_vl(j)
                assert (ost, ome, oopp, od) == (int(st[r]), int(lm[r]), int(lo[r]), int(plen[r])), (it, i, j)
            for j in range(leaves):
                r = j * B + i
                t.expand_backup_vl(j, w[r], v[r])
        s.prior_w.copy_(torch.from_numpy(w))
        s.value.copy_(torch.from_numpy(v))
        if it + 1 < iters:
            s.step()
        else:
            s.expand_backup()
    cnt, pi, q = (x.cpu().numpy() for x in s.root_policy())
    W, P = _root_W_P(s)
    for i, t in enumerate(trees):
        c, w_, p_ = t.root_stats()
        assert np.array_equal(cnt[i], c) and np.array_equal(W[i], w_) and np.array_equal(P[i], p_)
    s.check_errors()


@pytest.mark.parametrize("group", GROUPS)
@pytest.mark.parametrize("prefix", ["rev8_playout_s48_k2", "rev8_playout_s48_k4", "rev8_start_s240_k4"])
def test_virtual_loss_matches_golden(golden_mcts_vl, prefix, group):
    """golden = the Python definition run on the LIVE reference boards (tests/golden/mcts_vl.npz)"""
    g = golden_mcts_vl
    n_sims, salt, leaves = (int(v) for v in g[prefix + "_meta"])
    s, cnt, pi, q = _search_vl(g[prefix + "_me"], g[prefix + "_opp"], n_sims, salt, leaves, c_puct=float(g["c_puct"]),
                               group_lanes=group)
    assert np.array_equal(cnt, g[prefix + "_counts"])
    W, P = _root_W_P(s)
    assert np.array_equal(W, g[prefix + "_W"]) and np.array_equal(P, g[prefix + "_P"])


@pytest.mark.parametrize("leaves", [1, 4])
def test_path_depth_overflow_is_detected(leaves):
    """max_depth smaller than the trees grow: code 2 is flagged and raised, nothing is written past the path records"""
    from betazero_b200 import _lib, env, mcts
    from oracle import pyoracle as po

    me_h, opp_h = po.playout_boards(16, seed=4)
    pools = mcts.TreePools(16, 200, max_depth=2, n_leaves=leaves)
    s = mcts.BatchedMCTS(pools, mcts.HashEvaluator(0), use_graph=False)
    with pytest.raises(_lib.BzError, match="path depth"):
        s.search(env.to_device_u64(me_h), env.to_device_u64(opp_h), 200)
    assert int(pools.path_len.max().item()) <= 2


def test_virtual_loss_large_batch_stays_in_wave_mode():
    """>= 8192 trees with 4 leaves per iteration: still the wave kernels (a warp per tree); same answers"""
    from oracle import pyoracle as po

    B, S = 8192 + 3, 64
    me_h, opp_h = po.playout_boards(256, seed=9)
    me_h, opp_h = np.resize(me_h, B), np.resize(opp_h, B)
    s, cnt, _, _ = _search_vl(me_h, opp_h, S, 2, 4)
    r_cnt, _, _, _ = po.search_hash(me_h[:256], opp_h[:256], S, po.GAME_REVERSI, 8, 1.25, 2, leaves=4)
    assert np.array_equal(cnt[:256], r_cnt)
    assert np.array_equal(cnt[256:512], r_cnt) and np.array_equal(cnt[-3:], r_cnt[(B - 3) % 256:][:3])


@pytest.mark.parametrize("leaves", [1, 4])
def test_visit_counts_beyond_the_sqrt_table(leaves):
    """sqrt(n_node) comes from a 65536-entry table, and the straight-line scores are guarded by a vote that sends larger
    counts (and out-of-range operands) through the exact forms: a 4x4 search long enough for nodes with more than 65536
    visits (the small game tree saturates, most simulations end on known terminal nodes), bit for bit against the oracle."""
    from betazero_b200 import env, mcts
    from oracle import pyoracle as po

    B, S = 6, 70000
    me_h, opp_h = po.playout_boards(B, seed=3, size=4)
    me_h[0], opp_h[0] = po.grid_to_wire(po.OracleReversiBoard(size=4).board, 1)
    pools = mcts.TreePools(B, S, board_size=4, n_leaves=leaves, arena_units=400000)
    s = mcts.BatchedMCTS(pools, mcts.HashEvaluator(2))
    cnt, _, _ = s.search(env.to_device_u64(me_h), env.to_device_u64(opp_h), S)
    r_cnt, r_W, r_P, _ = po.search_hash(me_h, opp_h, S, po.GAME_REVERSI, 4, 1.25, 2, leaves=leaves)
    assert np.array_equal(cnt.cpu().numpy(), r_cnt)
    assert int(r_cnt.max()) > 65536
    W, P = _root_W_P(s)
    assert np.array_equal(W, r_W) and np.array_equal(P, r_P)
