"""bz_mcts_search_fused -- the whole search of a move in one kernel (tree islands + the MLP on CTA pairs) -- must build
exactly the trees of the per-iteration kernels (bz_mcts_select / bz_mlp_forward_pair* / bz_mcts_step), which are the
ones pinned to the oracle by test_gpu_mcts.py: same visit counts, W, P, node blocks and allocator state, bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _roots(B, seed, start=False):
    from betazero_b200 import env
    from oracle import pyoracle as po

    if start:
        me, opp, _ = env.reversi_init(B)
        return me, opp
    me_h, opp_h = po.playout_boards(B, seed=seed)  # midgame positions, including forced passes and finished games
    return env.to_device_u64(me_h), env.to_device_u64(opp_h)


def _search(model, me, opp, n_sims, one_launch, n_leaves=4, **kw):
    from betazero_b200 import mcts

    B = me.numel()
    s = mcts.BatchedMCTS(mcts.TreePools(B, n_sims, n_leaves=n_leaves), mcts.FusedNetEvaluator(model), use_graph=False,
                         one_launch=one_launch, **kw)
    assert s.one_launch is bool(one_launch)
    s.pools.arena.zero_()  # the padding words at the end of a node block are never written: make them comparable
    s.reset(me, opp)
    s.run(n_sims)
    torch.cuda.synchronize()
    s.check_errors()
    return s


def _same_trees(a, b):
    for x, y in zip(a.root_edges(), b.root_edges()):
        assert torch.equal(x, y)
    pa, pb = a.pools, b.pools
    for name in ("arena_used", "edge_count", "sim_count", "depth_sum", "root_meta"):
        assert torch.equal(getattr(pa, name), getattr(pb, name)), name
    used = pa.arena_used.long() * 8
    arena_a = pa.arena.view(pa.n_trees, -1)
    arena_b = pb.arena.view(pb.n_trees, -1)
    col = torch.arange(arena_a.shape[1], device=arena_a.device)[None, :]
    live = col < used[:, None]
    assert torch.equal(arena_a[live], arena_b[live])  # every node block of every tree


@pytest.mark.parametrize("B", [1, 13, 28, 29, 56, 57, 300, 4144, 4145, 9000])  # above 4144: chunks, one launch each
def test_one_launch_search_builds_the_same_trees(B):
    from betazero_b200 import net

    model = net.make_net("mlp", seed=B)
    me, opp = _roots(B, seed=100 + B)
    n_sims = 64 if B <= 300 else 32
    _same_trees(_search(model, me, opp, n_sims, True), _search(model, me, opp, n_sims, False))


@pytest.mark.parametrize("B,n_sims", [(1, 40), (29, 64), (300, 64), (4144, 16), (4500, 16)])
def test_one_launch_search_with_one_leaf_per_iteration_builds_the_same_trees(B, n_sims):
    """the sequential definition (one descent per tree and iteration, a full warp per tree) through search_fused_kernel<1>"""
    from betazero_b200 import net

    model = net.make_net("mlp", seed=B + 1)
    me, opp = _roots(B, seed=200 + B)
    _same_trees(_search(model, me, opp, n_sims, True, n_leaves=1), _search(model, me, opp, n_sims, False, n_leaves=1))


@pytest.mark.parametrize("B,n_sims", [(1, 40), (57, 64), (300, 96), (4144, 16), (4500, 16)])
def test_one_launch_search_with_two_leaves_per_iteration_builds_the_same_trees(B, n_sims):
    """two virtual-loss descents per tree and iteration (16 lanes each) through search_fused_kernel<2>"""
    from betazero_b200 import net

    model = net.make_net("mlp", seed=B + 2)
    me, opp = _roots(B, seed=300 + B)
    _same_trees(_search(model, me, opp, n_sims, True, n_leaves=2), _search(model, me, opp, n_sims, False, n_leaves=2))


def test_one_launch_two_leaves_headline_size_and_deep_trees():
    from betazero_b200 import net

    model = net.make_net("mlp", seed=0)
    me, opp = _roots(4096, 0, start=True)
    _same_trees(_search(model, me, opp, 400, True, n_leaves=2), _search(model, me, opp, 400, False, n_leaves=2))
    with torch.no_grad():  # peaked priors: paths deeper than the 16 entries a slot keeps in registers
        model.policy.weight.mul_(512)
        model.policy.bias.mul_(512)
    me, opp = _roots(256, 0, start=True)
    a, b = _search(model, me, opp, 600, True, n_leaves=2), _search(model, me, opp, 600, False, n_leaves=2)
    _same_trees(a, b)
    assert a.stats()["mean_depth"] > 18.0


def test_one_launch_one_leaf_headline_size_and_deep_trees():
    from betazero_b200 import net

    model = net.make_net("mlp", seed=0)
    me, opp = _roots(4096, 0, start=True)
    a, b = _search(model, me, opp, 200, True, n_leaves=1), _search(model, me, opp, 200, False, n_leaves=1)
    _same_trees(a, b)
    assert int(a.root_edges()[0][0].sum()) == 199  # the first iteration expands the root
    with torch.no_grad():  # peaked priors: paths deeper than the 32 entries a warp keeps in registers
        model.policy.weight.mul_(512)
        model.policy.bias.mul_(512)
    me, opp = _roots(256, 0, start=True)
    a, b = _search(model, me, opp, 600, True, n_leaves=1), _search(model, me, opp, 600, False, n_leaves=1)
    _same_trees(a, b)
    assert a.stats()["mean_depth"] > 20.0


def test_one_launch_headline_config_4096_trees_800_sims():
    from betazero_b200 import net

    model = net.make_net("mlp", seed=0)
    me, opp = _roots(4096, 0, start=True)
    a, b = _search(model, me, opp, 800, True), _search(model, me, opp, 800, False)
    _same_trees(a, b)
    assert int(a.root_edges()[0][0].sum()) == 796  # K = 4: the first iteration's four descents all end on the root


@pytest.mark.parametrize("B,n_sims", [(512, 400), (4300, 240)])  # 4300 trees: two chunks, each with its own part of the path array
def test_one_launch_deep_trees_paths_longer_than_a_lane_group(B, n_sims):
    """Peaked priors (policy head x 256) send the descents deep: path entries beyond depth 8 leave the lanes' registers
    and go through the path array in memory, in the one-launch search as in the per-iteration kernels."""
    from betazero_b200 import net

    model = net.make_net("mlp", seed=5)
    with torch.no_grad():
        model.policy.weight.mul_(256)
        model.policy.bias.mul_(256)
    me, opp = _roots(B, 0, start=True)
    a, b = _search(model, me, opp, n_sims, True), _search(model, me, opp, n_sims, False)
    _same_trees(a, b)
    assert a.stats()["mean_depth"] > 9.0


def test_one_launch_small_boards_and_single_iteration():
    from betazero_b200 import env, mcts, net

    model = net.make_net("mlp", seed=4)
    for size, n_sims in ((6, 40), (4, 24), (8, 4), (8, 8)):
        me, opp, _ = env.reversi_init(77, size)
        out = []
        for one in (True, False):
            s = mcts.BatchedMCTS(mcts.TreePools(77, n_sims, board_size=size, n_leaves=4), mcts.FusedNetEvaluator(model),
                                 use_graph=False, one_launch=one)
            s.reset(me, opp)
            s.run(n_sims)
            s.check_errors()
            out.append([x.clone() for x in s.root_edges()])
        for x, y in zip(*out):
            assert torch.equal(x, y), (size, n_sims)


def test_one_launch_with_root_noise_and_after_a_weight_update():
    from betazero_b200 import mcts, net

    model = net.make_net("mlp", seed=9)
    me, opp = _roots(200, 3)
    res = []
    for one in (True, False):
        s = mcts.BatchedMCTS(mcts.TreePools(200, 48, n_leaves=4), mcts.FusedNetEvaluator(model), use_graph=False,
                             one_launch=one, dirichlet_alpha=0.3, noise_seed=5)
        res.append(s.search(me, opp, 48)[0].clone())
    assert torch.equal(res[0], res[1])
    # the weight image is read at every launch: a refresh after an in-place update is seen by the next search
    ev = mcts.FusedNetEvaluator(model)
    s = mcts.BatchedMCTS(mcts.TreePools(200, 48, n_leaves=4), ev, one_launch=True)
    c0 = s.search(me, opp, 48)[0].clone()
    g = torch.Generator(device="cuda").manual_seed(1)
    with torch.no_grad():
        for prm in model.parameters():
            prm.add_(torch.randn(prm.shape, device="cuda", generator=g).to(prm.dtype) * 0.3)
    ev.refresh()
    c1 = s.search(me, opp, 48)[0].clone()
    ref = mcts.BatchedMCTS(mcts.TreePools(200, 48, n_leaves=4), mcts.FusedNetEvaluator(model), use_graph=False, one_launch=False)
    assert torch.equal(c1, ref.search(me, opp, 48)[0]) and not torch.equal(c0, c1)


def test_one_launch_is_refused_outside_its_shape():
    from betazero_b200 import _lib, mcts, net

    model = net.make_net("mlp", seed=1)
    big = mcts.BatchedMCTS(mcts.TreePools(4145, 8, n_leaves=4, arena_units=64), mcts.FusedNetEvaluator(model), use_graph=False)
    assert big.one_launch  # more than one launch's 4144 trees: searched in chunks
    for kw in (dict(n_leaves=8), dict(n_leaves=3), dict(n_leaves=4, group_lanes=8), dict(n_leaves=2, group_lanes=16)):
        s = mcts.BatchedMCTS(mcts.TreePools(64, 8, **kw), mcts.FusedNetEvaluator(model), use_graph=False)
        assert not s.one_launch
    with pytest.raises(RuntimeError):
        mcts.BatchedMCTS(mcts.TreePools(64, 9, n_leaves=3), mcts.FusedNetEvaluator(model), one_launch=True).one_launch
    assert not mcts.BatchedMCTS(mcts.TreePools(64, 8, n_leaves=4), mcts.HashEvaluator(1), use_graph=False).one_launch
    L = _lib.load()
    p = mcts.TreePools(64, 9, n_leaves=3, prior_mode=mcts.PRIOR_LOGITS_BF16, eval_stride=72)
    out = torch.zeros((192, 72), dtype=torch.bfloat16, device="cuda")
    model.prepare_inference()
    assert L.bz_mcts_search_fused(p._ref, _lib.dptr(model._image_pair), _lib.dptr(out), 2, _lib.stream_ptr()) == -1  # BZ_ERR_ARG


def test_selfplay_games_are_identical_with_the_one_launch_search():
    from betazero_b200 import mcts, net, selfplay

    model = net.make_net("mlp", seed=2)
    recs = []
    for one in (True, False):
        sp = selfplay.BatchedSelfPlay(96, 16, mcts.FusedNetEvaluator(model), board_size=6, temp_plies=6, seed=11, n_leaves=4,
                                      use_graph=False, one_launch=one)
        assert sp.mcts.one_launch is one
        for _ in range(40):
            sp.play_move()
        sp.mcts.check_errors()
        st = sp.stats()
        assert st["games"] >= 96 and st["dropped"] == 0
        r = sp.drain_replay()
        order = torch.argsort(r["game"] * 256 + r["ply"].long())  # finished games reserve their rows in any order
        recs.append((st, {k: v[order].cpu() for k, v in r.items()}))
    assert recs[0][0] == recs[1][0]
    for k in recs[0][1]:
        assert torch.equal(recs[0][1][k], recs[1][1][k]), k


@pytest.mark.parametrize("one", [True, False])
def test_search_host_returns_the_policy_and_move_of_a_fresh_search(one):
    """BatchedMCTS.search_host (host roots in, pinned pi + move out; with the one-launch search a replayed CUDA graph)
    against search() + best_action(), over several calls with different roots."""
    from betazero_b200 import mcts, net

    model = net.make_net("mlp", seed=6)
    B, n_sims = 200, 48
    s = mcts.BatchedMCTS(mcts.TreePools(B, n_sims, n_leaves=4), mcts.FusedNetEvaluator(model), use_graph=False, one_launch=one)
    ref = mcts.BatchedMCTS(mcts.TreePools(B, n_sims, n_leaves=4), mcts.FusedNetEvaluator(model), use_graph=False, one_launch=False)
    for seed in (1, 2, 3):
        me, opp = _roots(B, seed)
        h_pi, h_act = s.search_host(me.cpu(), opp.cpu(), n_sims)
        _, pi, _ = ref.search(me, opp, n_sims)
        act = ref.best_action()
        assert torch.equal(h_pi, pi.cpu()) and torch.equal(h_act, act.cpu())
        assert h_pi.is_pinned() and h_act.is_pinned()
