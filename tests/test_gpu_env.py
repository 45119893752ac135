"""GPU parity: CUDA env kernels (K1-K4, K6) through the C ABI vs the oracle and the golden
reference outputs.  Bit-exact (integer work)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


def _dev(a):
    from betazero_b200 import env

    return env.to_device_u64(a)


@pytest.mark.parametrize("size", [4, 6, 8])
def test_env_kernels_match_reference_golden(golden_env, size):
    from betazero_b200 import env

    g = golden_env
    me, opp = _dev(g[f"s{size}_me"]), _dev(g[f"s{size}_opp"])
    assert np.array_equal(env.to_host_u64(env.legal_mask(me, opp, size)), g[f"s{size}_mask"])
    assert np.array_equal(env.to_host_u64(env.legal_mask(opp, me, size)), g[f"s{size}_mask_opp"])
    over, win, cm, co = env.terminal(me, opp, size)
    assert np.array_equal(over.cpu().numpy(), g[f"s{size}_over"])
    assert np.array_equal(win.cpu().numpy(), g[f"s{size}_winner"])
    assert np.array_equal(cm.cpu().numpy(), g[f"s{size}_cnt_me"])
    assert np.array_equal(co.cpu().numpy(), g[f"s{size}_cnt_opp"])
    idx = torch.from_numpy(g[f"s{size}_succ_idx"].astype(np.int64)).cuda()
    act = torch.from_numpy(g[f"s{size}_succ_act"]).cuda()
    mo, oo, err = env.apply(me[idx].contiguous(), opp[idx].contiguous(), act, size)
    assert not err.any().item()
    assert np.array_equal(env.to_host_u64(mo), g[f"s{size}_succ_me"])
    assert np.array_equal(env.to_host_u64(oo), g[f"s{size}_succ_opp"])
    bidx = torch.from_numpy(g[f"s{size}_bad_idx"].astype(np.int64)).cuda()
    bact = torch.from_numpy(g[f"s{size}_bad_act"]).cuda()
    bm, bo = me[bidx].contiguous(), opp[bidx].contiguous()
    mo, oo, err = env.apply(bm, bo, bact, size)
    assert err.all().item()  # ValueError("Invalid move") in the reference
    assert torch.equal(mo, bm) and torch.equal(oo, bo)  # passed through unchanged


@pytest.mark.parametrize("size", [4, 6, 8])
def test_reference_episode_replay(golden_env, size):
    """a whole reference game (reversi_terminal.py loop) replayed ply by ply, passes included"""
    from betazero_b200 import env

    g = golden_env
    me, opp, act = g[f"ep{size}_me"], g[f"ep{size}_opp"], g[f"ep{size}_action"]
    m, o, err = env.apply(_dev(me), _dev(opp), torch.from_numpy(act).cuda(), size)
    assert not err.any().item()
    assert np.array_equal(env.to_host_u64(m)[:-1], me[1:]) and np.array_equal(env.to_host_u64(o)[:-1], opp[1:])
    over, win, cm, co = env.terminal(m, o, size)
    over = over.cpu().numpy()
    assert over[-1] == 1 and not over[:-1].any()
    w, c1, c2 = g[f"ep{size}_score"]
    last_player = -int(g[f"ep{size}_player"][-1])  # mover of the final position
    assert int(win[-1].item()) == w * last_player
    got = (int(cm[-1].item()), int(co[-1].item()))
    assert got == ((c1, c2) if last_player == 1 else (c2, c1))
    i_me, i_opp, i_pl = env.reversi_init(3, size)
    assert env.to_host_u64(i_me)[0] == me[0] and env.to_host_u64(i_opp)[0] == opp[0] and int(i_pl[0].item()) == 1


def test_config2_one_million_boards_vs_oracle():
    """BASELINE config 2: legal mask + next state on 1M random synthetic boards, bit-exact vs
    the C restatement of the reference ray walk (oracle.c), plus 64K reachable boards."""
    from betazero_b200 import env
    from oracle import pyoracle as po

    n = 1 << 20
    me_h, opp_h = po.synthetic_boards(n, seed=0)
    pm, pop = po.playout_boards(1 << 14, seed=0)
    me_h, opp_h = np.concatenate([me_h, pm]), np.concatenate([opp_h, pop])
    me, opp = _dev(me_h), _dev(opp_h)
    mask = env.legal_mask(me, opp)
    ref_mask = po.legal_mask(me_h, opp_h)
    assert np.array_equal(env.to_host_u64(mask), ref_mask)
    # fused step: lowest legal move (or pass)
    k, act, m2, o2 = env.step_first_legal(me, opp)
    assert np.array_equal(env.to_host_u64(k), ref_mask)
    lsb = np.maximum(ref_mask & (~ref_mask + np.uint64(1)), np.uint64(1)).astype(np.float64)  # exact: a power of two
    low = np.where(ref_mask != 0, np.log2(lsb).astype(np.int64), 64)
    assert np.array_equal(act.cpu().numpy().astype(np.int64), low)
    rm, ro, rerr = po.apply(me_h, opp_h, low.astype(np.uint8))
    assert not rerr.any()
    assert np.array_equal(env.to_host_u64(m2), rm) and np.array_equal(env.to_host_u64(o2), ro)
    # K2 on its own gives the same successors; K3 vs oracle
    m3, o3, err = env.apply(me, opp, act)
    assert not err.any().item() and torch.equal(m3, m2) and torch.equal(o3, o2)
    over, win, cm, co = env.terminal(me, opp)
    r_over, r_win, r_cm, r_co = po.terminal(me_h, opp_h)
    assert np.array_equal(over.cpu().numpy(), r_over) and np.array_equal(win.cpu().numpy(), r_win)
    assert np.array_equal(cm.cpu().numpy(), r_cm) and np.array_equal(co.cpu().numpy(), r_co)


def test_random_legal_and_illegal_actions_vs_oracle():
    from betazero_b200 import env
    from oracle import pyoracle as po

    n = 200_000
    me_h, opp_h = po.synthetic_boards(n, seed=3)
    rng = np.random.default_rng(4)
    act_h = rng.integers(0, 66, size=n).astype(np.uint8)  # includes pass (64) and an out-of-range id (65)
    mo, oo, err = env.apply(_dev(me_h), _dev(opp_h), torch.from_numpy(act_h).cuda())
    rm, ro, rerr = po.apply(me_h, opp_h, act_h)
    assert np.array_equal(err.cpu().numpy(), rerr)
    assert np.array_equal(env.to_host_u64(mo), rm) and np.array_equal(env.to_host_u64(oo), ro)
    assert 0 < rerr.sum() < n


@pytest.mark.parametrize("n", [0, 1, 2, 3, 255, 257])
def test_ragged_sizes_and_unaligned_views(n):
    from betazero_b200 import env
    from oracle import pyoracle as po

    me_h, opp_h = po.synthetic_boards(n + 1, seed=n)
    me, opp = _dev(me_h), _dev(opp_h)
    # [1:] views are 8-byte (not 16-byte) aligned: exercises the scalar path
    for (a, b, ah, bh) in ((me[:n], opp[:n], me_h[:n], opp_h[:n]), (me[1:], opp[1:], me_h[1:], opp_h[1:])):
        got = env.to_host_u64(env.legal_mask(a, b)) if a.numel() else np.zeros(0, np.uint64)
        assert np.array_equal(got, po.legal_mask(ah, bh) if ah.size else np.zeros(0, np.uint64))
        if a.numel():
            k, act, m2, o2 = env.step_first_legal(a, b)
            rm, ro, rerr = po.apply(ah, bh, act.cpu().numpy())
            assert not rerr.any() and np.array_equal(env.to_host_u64(m2), rm) and np.array_equal(env.to_host_u64(o2), ro)


def test_full_scale_properties():
    """size-independent checks at 2^24 boards (beyond what the CPU oracle checks quickly)"""
    from betazero_b200 import env

    n = 1 << 24
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randint(-(2 ** 63), 2 ** 63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    b = torch.randint(-(2 ** 63), 2 ** 63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    c = torch.randint(-(2 ** 63), 2 ** 63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    me, opp = a & b, ~a & c  # disjoint, ~25% / ~25% occupancy
    mask = env.legal_mask(me, opp)
    assert not (mask & (me | opp)).any().item()  # legal cells are empty
    k, act, m2, o2 = env.step_first_legal(me, opp)
    assert torch.equal(k, mask)
    moved = act != 64
    assert torch.equal(moved, mask != 0)
    assert not (m2 & o2).any().item()
    # discs never vanish: occupied' = occupied + the placed cell
    placed = torch.where(moved, torch.ones_like(me) << act.to(torch.int64).clamp(max=63), torch.zeros_like(me))
    assert torch.equal(m2 | o2, me | opp | placed)
    # the mover's discs only grow (now seen as opp'), at least one flip when a move was made
    assert torch.equal(o2 & me, me)
    _, _, cm, co = env.terminal(me, opp)
    _, _, cm2, co2 = env.terminal(m2, o2)
    gained = co2.to(torch.int32) - cm.to(torch.int32)
    assert (gained[moved] >= 2).all().item() and (gained[~moved] == 0).all().item()
    assert torch.equal(cm.to(torch.int32) + co.to(torch.int32) + moved.to(torch.int32), cm2.to(torch.int32) + co2.to(torch.int32))
    # pass twice = identity
    p = torch.full((n,), 64, dtype=torch.uint8, device="cuda")
    nm = mask == 0
    m3, o3, e3 = env.apply(me, opp, p)
    assert torch.equal(e3 == 0, nm)
    assert torch.equal(m3[nm], opp[nm]) and torch.equal(o3[nm], me[nm])


def test_planes_canonical_form():
    from betazero_b200 import env
    from oracle import pyoracle as po

    me_h, opp_h = po.synthetic_boards(4097, seed=9)
    pl = env.planes(_dev(me_h), _dev(opp_h))
    assert pl.shape == (4097, 2, 8, 8) and pl.dtype == torch.bfloat16
    got = pl.float().cpu().numpy()
    bits = np.arange(64, dtype=np.uint64)
    exp_me = ((me_h[:, None] >> bits) & np.uint64(1)).astype(np.float32).reshape(-1, 8, 8)
    exp_opp = ((opp_h[:, None] >> bits) & np.uint64(1)).astype(np.float32).reshape(-1, 8, 8)
    assert np.array_equal(got[:, 0], exp_me) and np.array_equal(got[:, 1], exp_opp)
    # players.py:85: symbol * board == plane0 - plane1 for the mover
    g0 = po.wire_to_grid(me_h[0], opp_h[0])
    assert np.array_equal(got[0, 0] - got[0, 1], g0.astype(np.float32))


def test_ttt_kernels_match_reference_golden(golden_ttt):
    from betazero_b200 import env
    from oracle import pyoracle as po

    g = golden_ttt
    x, o = env.to_device_u16(g["x"]), env.to_device_u16(g["o"])
    assert np.array_equal(env.to_host_u16(env.ttt_legal_mask(x, o)), g["mask"])
    over, win = env.ttt_terminal(x, o)
    assert np.array_equal(over.cpu().numpy(), g["over"]) and np.array_equal(win.cpu().numpy(), g["winner"])
    rng = np.random.default_rng(0)
    act = rng.integers(0, 10, size=g["x"].size).astype(np.uint8)
    pl = rng.choice(np.array([1, -1], dtype=np.int8), size=g["x"].size)
    xo, oo, err = env.ttt_apply(x, o, torch.from_numpy(act).cuda(), torch.from_numpy(pl).cuda())
    rx, ro, rerr = po.ttt_apply(g["x"], g["o"], act, pl)
    assert np.array_equal(env.to_host_u16(xo), rx) and np.array_equal(env.to_host_u16(oo), ro)
    assert np.array_equal(err.cpu().numpy(), rerr)
    # the reference's CSV golden: canonical state + one-hot action -> next state
    st, ac = g["csv_state"], g["csv_action"]
    xs = np.array([sum(1 << i for i in range(9) if s[i] == 1) for s in st], np.uint16)
    os_ = np.array([sum(1 << i for i in range(9) if s[i] == -1) for s in st], np.uint16)
    k = ac.argmax(1).astype(np.uint8)
    xo, oo, err = env.ttt_apply(env.to_device_u16(xs), env.to_device_u16(os_), torch.from_numpy(k).cuda(),
                                torch.ones(len(k), dtype=torch.int8, device="cuda"))
    assert not err.any().item()
    assert np.array_equal(env.to_host_u16(xo), xs | (np.uint16(1) << k.astype(np.uint16)))
    assert np.array_equal(env.to_host_u16(oo), os_)
