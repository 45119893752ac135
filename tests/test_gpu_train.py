"""GPU: SURVEY.md 8f rows 2 and config 5 -- checkpoint round trip (the reference saves its model at the end of training,
SL/train.py:204-214), a player built from a model FILE (AIPlayer(path_to_model, symbol), players.py:77-98), and one full
iteration of the AlphaZero loop (self-play -> replay gather -> training -> weight broadcast) through betazero_b200.loop."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


def _records(n, seed):
    """n plausible (board, pi, z) training records without running a search"""
    from betazero_b200 import env
    from oracle import pyoracle as po

    me_h, opp_h = po.playout_boards(n, seed=seed)
    g = torch.Generator(device="cuda").manual_seed(seed)
    pi = torch.rand((n, 65), device="cuda", generator=g)
    pi /= pi.sum(1, keepdim=True)
    z = torch.randint(-1, 2, (n,), device="cuda", generator=g).to(torch.int8)
    return {"me": env.to_device_u64(me_h), "opp": env.to_device_u64(opp_h), "pi": pi, "z": z}


def test_checkpoint_round_trip_restores_model_optimizer_and_iteration(tmp_path):
    from betazero_b200 import net, train

    rp = _records(512, seed=3)
    idx = torch.arange(512, device="cuda")
    planes, pi, z = train.make_batch(rp, idx)
    model = net.make_net("mlp", seed=0, dtype=torch.float32)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    for _ in range(3):
        train.train_step(model, opt, planes, pi, z)
    path = str(tmp_path / "ck.pt")
    train.save_checkpoint(path, model, opt, iteration=7, extra={"note": "round trip"})

    raw = torch.load(path, map_location="cpu", weights_only=True)  # no pickled code: loads in the safe mode
    assert raw["iteration"] == 7 and raw["extra"] == {"note": "round trip"}
    assert raw["arch"] == {"kind": "mlp", "in_features": 128, "hidden": 256, "n_actions": 65}

    model2 = net.make_net("mlp", seed=99, dtype=torch.float32)  # different init: everything must come from the file
    opt2 = torch.optim.Adam(model2.parameters(), lr=5e-2)
    assert train.load_checkpoint(path, model2, opt2) == 7
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k
    sd, sd2 = opt.state_dict(), opt2.state_dict()
    assert sd["param_groups"] == sd2["param_groups"]  # lr 1e-3 came back, not 5e-2
    for pid in sd["state"]:
        for key in ("step", "exp_avg", "exp_avg_sq"):
            assert torch.equal(sd["state"][pid][key].cpu(), sd2["state"][pid][key].cpu()), (pid, key)
    with torch.no_grad():
        la, va = model(planes)
        lb, vb = model2(planes)
    assert torch.equal(la, lb) and torch.equal(va, vb)
    # training continues identically from the restored state
    a = train.train_step(model, opt, planes, pi, z)
    b = train.train_step(model2, opt2, planes, pi, z)
    assert a["loss"] == pytest.approx(b["loss"], rel=1e-6)
    for pa, pb in zip(model.parameters(), model2.parameters()):
        torch.testing.assert_close(pa, pb, rtol=1e-6, atol=1e-8)


def test_load_checkpoint_refreshes_the_inference_buffers_in_place(tmp_path):
    """a bf16 inference net that already serves a search: loading a checkpoint must update the fused head / weight
    image at the same addresses (captured graphs keep their pointers)"""
    from betazero_b200 import net, train

    src = net.make_net("mlp", seed=5)
    path = str(tmp_path / "m.pt")
    train.save_checkpoint(path, src)
    dst = net.make_net("mlp", seed=6)
    x = torch.zeros((300, 2, 8, 8), dtype=torch.bfloat16, device="cuda")
    x[:, 0, 3, 3] = 1
    before = dst.forward_raw(x).clone()
    ptr = dst._image_pair.data_ptr()
    train.load_checkpoint(path, dst)
    assert dst._image_pair.data_ptr() == ptr
    after = dst.forward_raw(x)
    assert torch.equal(after, src.forward_raw(x)) and not torch.equal(after, before)


def test_aiplayer_is_built_from_a_model_path(tmp_path):
    """the reference's AIPlayer(path_to_model, symbol) convention (players.py:77-81): same moves as the net it was saved
    from (argmax over the LEGAL moves, :92-98); with n_sims > 0 the same file drives the MCTS player"""
    from betazero_b200 import net, train
    from betazero_b200.boards import ReversiBoard
    from betazero_b200.players import AIPlayer, GreedyNetPlayer

    model = net.make_net("mlp", seed=8)
    path = str(tmp_path / "reversi_model.pt")
    train.save_checkpoint(path, model)
    ai, greedy = AIPlayer(path, 1), GreedyNetPlayer(1, model)
    ai_o, greedy_o = AIPlayer(path, -1), GreedyNetPlayer(-1, model)
    b, player = ReversiBoard(size=8), 1
    for _ in range(12):
        moves = b.generate_possible_moves(player)
        if not moves:
            break
        mv = (ai if player == 1 else ai_o).get_move(b)
        assert mv == (greedy if player == 1 else greedy_o).get_move(b) and mv in moves
        b = b.make_move(*mv, player)
        player = -player
    searcher = AIPlayer(path, 1, n_sims=32, n_leaves=4)
    b0 = ReversiBoard(size=8)
    assert searcher.get_move(b0) in b0.generate_possible_moves(1)
    for bad in ({"model": model.state_dict()},):  # a file without an architecture record is refused, not guessed
        p2 = str(tmp_path / "bare.pt")
        torch.save(bad, p2)
        with pytest.raises(ValueError, match="architecture"):
            AIPlayer(p2, 1)


def test_the_loop_runs_with_kept_trees():
    """--reuse: self-play of the iteration loop continues every search on the subtree of the move played"""
    from betazero_b200 import loop

    args = loop.default_args(games=48, sims=16, leaves=4, plies=40, size=6, train_steps=2, batch=128, temp_plies=30, reuse=True)
    st = loop.LoopState(args, rank=0, world=1)
    assert st.sp.reuse and st.sp.pools.scratch is not None
    line = loop.run_iteration(st, 0)
    assert line["games_finished_local"] >= 48 and line["records_dropped_local"] == 0 and line["train_steps"] == 2
    assert int(st.sp.pools.inherited.max()) > 0  # some root of the last search started with kept visits


@pytest.mark.parametrize("leaves", [1, 4])
def test_one_iteration_of_the_alphazero_loop(tmp_path, leaves):
    """BASELINE configs[4] on one GPU, small: full games with slot recycling, replay drain + (single-rank) gather,
    training steps on augmented batches, weight publication into the search's captured graph, checkpoint + resume"""
    from betazero_b200 import loop, train

    ck = str(tmp_path / "loop.pt")
    args = loop.default_args(games=48, sims=16, leaves=leaves, plies=40, size=6, train_steps=3, batch=256, lr=1e-3,
                             temp_plies=30, checkpoint=ck)
    st = loop.LoopState(args, rank=0, world=1)
    w0 = [p.detach().clone() for p in st.model.parameters()]
    image_ptr = st.model._image_pair.data_ptr()
    line = loop.run_iteration(st, 0)
    assert line["games_finished_local"] >= 48 and line["records_dropped_local"] == 0
    assert line["replay_records_gathered"] == line["local_records"] > 48 * 20
    assert line["train_steps"] == 3 and all(np.isfinite(line["loss_first_last"]))
    assert set(line["ms"]) == {"selfplay", "replay_gather", "train", "weight_broadcast"} and line["ms"]["selfplay"] > 0
    assert line["broadcast_bytes"] == 0  # one rank: nothing to broadcast
    # the trained weights reached the inference net, in place
    assert any(not torch.equal(a, b) for a, b in zip(w0, st.model.parameters()))
    assert st.model._image_pair.data_ptr() == image_ptr
    for m, i in zip(st.master.parameters(), st.model.parameters()):
        assert torch.equal(m.to(torch.bfloat16), i)
    # second iteration plays with the new net through the SAME captured graph
    line2 = loop.run_iteration(st, 1)
    assert line2["replay_records_gathered"] > 0 and line2["checkpoint"] == ck
    # resume: a new process state picks up weights, Adam moments and the iteration counter
    st2 = loop.LoopState(loop.default_args(games=48, sims=16, leaves=leaves, plies=2, size=6, resume=ck), rank=0, world=1)
    assert st2.first_iteration == 2
    for a, b in zip(st.master.parameters(), st2.master.parameters()):
        assert torch.equal(a, b)
    for a, b in zip(st.model.parameters(), st2.model.parameters()):
        assert torch.equal(a, b)
    assert train.load_checkpoint(ck, st2.master) == 2


def _dense(me, opp, pi, size):
    """(state [size, size] of +1 / -1 / 0, action [size, size], pass prior) of one record, like the reference's State / Action"""
    st = torch.zeros((8, 8))
    for b in range(64):
        if (me >> b) & 1:
            st[b >> 3, b & 7] = 1.0
        elif (opp >> b) & 1:
            st[b >> 3, b & 7] = -1.0
    return st[:size, :size].clone(), pi[:64].reshape(8, 8)[:size, :size].clone(), float(pi[64])


@pytest.mark.parametrize("size", [8, 6])
@pytest.mark.parametrize("reference_list", [False, True])
def test_expand_with_transforms_keeps_the_first_of_every_distinct_pair_like_the_reference(size, reference_list):
    """SL/train.py:23-50: every record under the eight lambdas, a (state string, action string) set keeps the first
    occurrence.  Restated here on dense tensors with the reference's own lambdas; the engine must keep the same
    (record, transform) pairs in the same order, with the same contents."""
    from betazero_b200 import env, train
    from oracle import pyoracle as po

    rng = np.random.default_rng(size)
    n = 60
    me_h, opp_h = po.playout_boards(n, seed=5, size=size)
    start = po.grid_to_wire(po.OracleReversiBoard(size=size).board, 1)
    me_h[:6], opp_h[:6] = start[0], start[1]  # symmetric positions: transforms coincide
    me_h[10:14], opp_h[10:14] = me_h[20:24], opp_h[20:24]  # repeated records
    pi_h = np.zeros((n, 65), np.float32)
    for i in range(n):
        cells = [r * 8 + c for r in range(size) for c in range(size) if not ((int(me_h[i]) | int(opp_h[i])) >> (r * 8 + c)) & 1]
        pick = rng.choice(cells, size=min(3, len(cells)), replace=False) if cells else []
        for a in pick:
            pi_h[i, a] = rng.integers(1, 5) / 8.0
        pi_h[i, 64] = 0.125 if i % 7 == 0 else 0.0
    pi_h[:6] = 0
    pi_h[:6, 64] = 1.0  # fully symmetric records: all eight transforms are one pair
    pi_h[10:14] = pi_h[20:24]
    lambdas = [lambda x: x, lambda x: x.flip(dims=[0]), lambda x: x.flip(dims=[1]), lambda x: x.rot90(1, [0, 1]),
               lambda x: x.rot90(2, [0, 1]), lambda x: x.rot90(3, [0, 1]), lambda x: x.t(),
               (lambda x: x.flip(dims=[0]).t()) if reference_list else (lambda x: x.flip(dims=[0, 1]).t())]
    seen, expect = set(), []
    for i in range(n):
        st, ac, ps = _dense(int(me_h[i]), int(opp_h[i]), torch.from_numpy(pi_h[i]), size)
        for k, f in enumerate(lambdas):
            ts, ta = f(st), f(ac)
            key = (",".join(map(str, ts.reshape(-1).tolist())), ",".join(map(str, ta.reshape(-1).tolist())), ps)
            if key not in seen:
                seen.add(key)
                expect.append((i, ts, ta))
    me, opp = env.to_device_u64(me_h), env.to_device_u64(opp_h)
    pi = torch.from_numpy(pi_h).cuda()
    z = torch.arange(n, dtype=torch.int8, device="cuda")
    tf = train.REFERENCE_TRANSFORMS if reference_list else tuple(range(8))
    me8, opp8, pi8, z8, src = train.expand_with_transforms(me, opp, pi, z, size=size, transforms=tf)
    assert src.cpu().tolist() == [e[0] for e in expect]
    assert torch.equal(z8.cpu(), z.cpu()[src.cpu()])
    mh, oh, ph = env.to_host_u64(me8), env.to_host_u64(opp8), pi8.cpu()
    for j, (i, ts, ta) in enumerate(expect):
        st, ac, ps = _dense(int(mh[j]), int(oh[j]), ph[j], size)
        assert torch.equal(st, ts) and torch.equal(ac, ta) and ps == float(pi_h[i, 64]), (j, i)
    if reference_list:
        assert len(expect) <= 7 * n  # the reference's 8th lambda never adds a record
    full = train.expand_with_transforms(me, opp, pi, size=size, dedup=False)
    assert full[0].numel() == 8 * n and len(expect) < 8 * n
