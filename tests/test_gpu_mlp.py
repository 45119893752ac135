"""GPU: the fused tcgen05 MLP kernels (bz_mlp_forward_pair / _pair2) against the PyTorch module they replace.
Floating point (bf16 inputs, fp32 accumulate, bf16 rounding after every layer), so the comparison
is tolerance based: the two differ only in fp32 accumulation order."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

ATOL, RTOL = 3e-2, 3e-2  # bf16 has 8 mantissa bits; logits are O(1)


def _planes(B, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randint(0, 3, (B, 64), device="cuda", generator=g)
    return torch.stack([(a == 1), (a == 2)], dim=1).reshape(B, 2, 8, 8).to(torch.bfloat16)


# "pair": bz_mlp_forward_pair (cta_group::2, weights resident); "pair2": two ping-ponged tiles per pair; True: by batch size
@pytest.mark.parametrize("variant", ["pair", "pair2", True])
@pytest.mark.parametrize("B", [1, 7, 64, 65, 128, 129, 257, 1000, 4096])
def test_fused_mlp_matches_torch_module(B, variant):
    from betazero_b200 import net

    m = net.make_net("mlp", seed=5)
    with torch.no_grad():  # larger weights than the default init so every layer matters
        for prm in m.parameters():
            prm.mul_(1.7)
    m.prepare_inference()
    x = _planes(B, B)
    assert m.fused_kernel_ok(x)
    got = m.forward_raw(x, fused=variant).float()
    ref = m.forward_raw(x, fused=False).float()
    torch.cuda.synchronize()
    assert got.shape == (B, 72)
    np.testing.assert_allclose(got[:, :66].cpu().numpy(), ref[:, :66].cpu().numpy(), atol=ATOL, rtol=RTOL)
    assert (got[:, 66:] == 0).all()  # zero padding columns
    with torch.no_grad():
        logits, v = m(x)
    np.testing.assert_allclose(got[:, :65].cpu().numpy(), logits.float().cpu().numpy(), atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(torch.tanh(got[:, 65]).cpu().numpy(), v.float().cpu().numpy(), atol=ATOL)
    # an fp64 reference on the same bf16 weights bounds both implementations
    with torch.no_grad():
        h = x.reshape(B, -1).double()
        for fc in (m.fc1, m.fc2, m.fc3):
            h = torch.relu(h @ fc.weight.double().t() + fc.bias.double()).to(torch.bfloat16).double()
        exact = h @ m.policy.weight.double().t() + m.policy.bias.double()
    assert (got[:, :65].double() - exact).abs().max().item() < 4e-2


def test_fused_mlp_is_deterministic_and_row_independent():
    from betazero_b200 import net

    m = net.make_net("mlp", seed=1)
    x = _planes(300, 3)
    a = m.forward_raw(x, fused="pair").clone()
    b = m.forward_raw(x, fused="pair")
    assert torch.equal(a, b)
    c = m.forward_raw(x[37:38].contiguous(), fused="pair")
    assert torch.equal(c[0], a[37])  # a row's result does not depend on its batch
    assert torch.equal(a, m.forward_raw(x, fused="pair2"))  # same accumulation order (K ascending, fp32 in TMEM)
    assert torch.equal(c[0], m.forward_raw(x[37:38].contiguous(), fused="pair2")[0])


def test_search_with_fused_mlp_kernel_agrees_with_library_gemms():
    from betazero_b200 import env, mcts, net
    from oracle import pyoracle as po

    B, n_sims = 256, 48
    me_h, opp_h = po.playout_boards(B, seed=8)
    model = net.make_net("mlp", seed=3)
    res = []
    for fused in (True, False):
        ev = mcts.FusedNetEvaluator(model, use_kernel=fused)
        pools = mcts.TreePools(B, n_sims)
        s = mcts.BatchedMCTS(pools, ev, use_graph=True, graph_unroll=8, one_launch=False)  # the per-iteration path
        cnt, pi, q = s.search(env.to_device_u64(me_h), env.to_device_u64(opp_h), n_sims)
        _, _, P = s.root_edges()
        res.append((cnt.cpu().numpy(), P.cpu().numpy()))
    np.testing.assert_allclose(res[0][1], res[1][1], atol=2e-2)
    live = po.terminal(me_h, opp_h)[0] == 0
    assert (res[0][0].sum(1)[live] == n_sims - 1).all()
    assert (res[0][0].argmax(1) == res[1][0].argmax(1))[live].mean() > 0.8


def test_programmatic_dependent_launch_changes_nothing_but_time():
    """FusedNetEvaluator(pdl=...): the step kernel and the fused MLP kernel overlap prologue/tail; results identical"""
    from betazero_b200 import _lib, env, mcts, net
    from oracle import pyoracle as po

    B, n_sims = 512, 40
    me_h, opp_h = po.playout_boards(B, seed=21)
    model = net.make_net("mlp", seed=7)
    out = []
    try:
        for pdl in (False, True):
            pools = mcts.TreePools(B, n_sims)
            s = mcts.BatchedMCTS(pools, mcts.FusedNetEvaluator(model, use_kernel=True, pdl=pdl), use_graph=True, graph_unroll=8,
                                 one_launch=False)  # the per-iteration path is the one that chains two kernels
            cnt, pi, q = s.search(env.to_device_u64(me_h), env.to_device_u64(opp_h), n_sims)
            out.append((cnt.clone(), q.clone()))
    finally:
        _lib.set_pdl(False)
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])


def test_pdl_flag_does_not_leak_into_searches_without_a_kernel_between_steps():
    """A search whose evaluator launches nothing between two step kernels must not inherit the launch
    attribute from an earlier FusedNetEvaluator search (adjacent step kernels would overlap)."""
    from betazero_b200 import _lib, env, mcts, net
    from oracle import pyoracle as po

    B, n_sims = 256, 64
    me_h, opp_h = po.playout_boards(B, seed=5)
    me, opp = env.to_device_u64(me_h), env.to_device_u64(opp_h)
    model = net.make_net("mlp", seed=7)
    s0 = mcts.BatchedMCTS(mcts.TreePools(B, n_sims), mcts.FusedNetEvaluator(model), graph_unroll=8, one_launch=False)
    s0.search(me, opp, n_sims)  # the per-iteration path: MLP kernel and step kernel chained with the launch attribute
    assert _lib._pdl_state is True

    class Static:  # fixed logits, no kernel of its own
        prior_mode, stride = mcts.PRIOR_LOGITS_BF16, 72

        def bind(self, pools):
            g = torch.Generator(device="cuda").manual_seed(3)
            self.out = torch.randn((B, 72), device="cuda", generator=g).to(torch.bfloat16)
            return self.out, torch.zeros(1, device="cuda")

        def __call__(self, pools):
            pass

    res = []
    for graph in (True, False):
        s1 = mcts.BatchedMCTS(mcts.TreePools(B, n_sims), Static(), use_graph=graph, graph_unroll=8)
        res.append(s1.search(me, opp, n_sims)[0].clone())
        assert _lib._pdl_state is False
    assert torch.equal(res[0], res[1])
    _lib.set_pdl(False)


def test_pair_kernels_are_race_free_under_repetition():
    """The pair kernels synchronise two CTAs per tile with mbarriers in each other's shared memory: hammer them (with and
    without the launch attribute for programmatic dependent launch, back to back on one stream) and require bit-identical
    results every time -- a missed ordering between the epilogue stores and the peer-issued MMAs would show up here."""
    from betazero_b200 import _lib, net

    m = net.make_net("mlp", seed=11)
    m.prepare_inference()
    try:
        for B, mode in ((4096, "pair"), (16384, "pair2"), (777, "pair"), (9999, "pair2")):
            x = _planes(B, B + 1)
            ref = m.forward_raw(x, fused="pair2" if mode == "pair" else "pair").clone()  # the OTHER kernel: bit-identical
            out = torch.empty_like(ref)
            for pdl in (False, True):
                _lib.set_pdl(pdl)
                bad = 0
                for it in range(200):
                    out.fill_(0)
                    m.forward_raw(x, out=out, fused=mode)
                    if it % 20 == 19:
                        bad += int(not torch.equal(out, ref))
                torch.cuda.synchronize()
                assert bad == 0 and torch.equal(out, ref), (B, mode, pdl)
    finally:
        _lib.set_pdl(False)


@pytest.mark.parametrize("use_kernel", [None, False])
@pytest.mark.parametrize("leaves", [1, 4])
def test_refresh_after_weight_update_reaches_the_captured_graph(use_kernel, leaves):
    """The [evaluate -> step] CUDA graph has the device pointers of the net's inference buffers (weight image, fused
    head) baked in.  After a weight update + ``evaluator.refresh()`` -- what loop.py does every iteration -- a graph
    replay must use the NEW weights: the buffers keep their addresses and are refreshed in place."""
    from betazero_b200 import env, mcts, net
    from oracle import pyoracle as po

    B, n_sims = 256, 72
    me_h, opp_h = po.playout_boards(B, seed=19)
    me, opp = env.to_device_u64(me_h), env.to_device_u64(opp_h)
    model = net.make_net("mlp", seed=3)
    ev = mcts.FusedNetEvaluator(model, use_kernel=use_kernel)
    s = mcts.BatchedMCTS(mcts.TreePools(B, n_sims, n_leaves=leaves), ev, use_graph=True, graph_unroll=4, one_launch=False)
    cnt0 = s.search(me, opp, n_sims)[0].clone()
    assert s._graph is not None
    ptrs = (model._head_w.data_ptr(), model._head_b.data_ptr(),
            model._image_pair.data_ptr() if model._image_pair is not None else 0)
    g = torch.Generator(device="cuda").manual_seed(1)
    with torch.no_grad():  # a "training step": every parameter changes in place
        for prm in model.parameters():
            prm.add_(torch.randn(prm.shape, device="cuda", generator=g).to(prm.dtype) * 0.3)
    ev.refresh()
    assert ptrs == (model._head_w.data_ptr(), model._head_b.data_ptr(),
                    model._image_pair.data_ptr() if model._image_pair is not None else 0)
    cnt1 = s.search(me, opp, n_sims)[0].clone()  # replays the graph captured BEFORE the update
    fresh = mcts.BatchedMCTS(mcts.TreePools(B, n_sims, n_leaves=leaves), mcts.FusedNetEvaluator(model, use_kernel=use_kernel),
                             use_graph=False, one_launch=False)
    cnt2 = fresh.search(me, opp, n_sims)[0]
    assert torch.equal(cnt1, cnt2)  # the graph saw the new weights
    assert not torch.equal(cnt0, cnt1)  # and they really differ from the old ones
