"""Smoke-sized runs of every kernel family for compute-sanitizer, or -- where the sanitizer is not available -- for a
library built with -DBZ_BOUNDS_CHECK (profiles/sanitize.sh): every pool index the tree / self-play kernels form is checked
on the device and the violation record is read back at the end.  Each stage also checks its own result against the oracle
/ the other implementation, so a clean run is also a correct run."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from betazero_b200 import env, mcts, net, selfplay  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

stage = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.cuda.set_device(0)

if stage in ("all", "tree"):
    me_h, opp_h = po.playout_boards(48, seed=1)
    me, opp = env.to_device_u64(me_h), env.to_device_u64(opp_h)
    for leaves, group in ((1, 32), (1, 8), (2, 0), (4, 0), (3, 0)):  # step_kernel<32>, <8>, step_wave_kernel<16>, <8>, step_kernel<VL>
        pools = mcts.TreePools(48, 48, n_leaves=leaves, group_lanes=group)
        s = mcts.BatchedMCTS(pools, mcts.HashEvaluator(5), use_graph=False)
        cnt = s.search(me, opp, 48)[0].cpu().numpy()
        ref = po.search_hash(me_h, opp_h, 48, salt=5, leaves=leaves)[0]
        assert np.array_equal(cnt, ref), (leaves, group)
    print("tree kernels ok")

if stage in ("all", "selfplay"):
    model = net.make_net("mlp", seed=2)
    sp = selfplay.BatchedSelfPlay(40, 16, mcts.FusedNetEvaluator(model), board_size=4, temp_plies=4, seed=5, use_graph=False,
                                  n_leaves=4)
    for _ in range(14):
        sp.play_move()
    sp.mcts.check_errors()
    st = sp.stats()
    assert st["games"] >= 30 and st["dropped"] == 0, st
    print("selfplay_advance + wave kernels + mlp_pair (PDL) ok", st)

if stage in ("all", "one_launch"):
    model = net.make_net("mlp", seed=6)
    for B, sims in ((57, 48), (300, 32)):
        me_h, opp_h = po.playout_boards(B, seed=B)
        me, opp = env.to_device_u64(me_h), env.to_device_u64(opp_h)
        res = []
        for one in (True, False):
            s = mcts.BatchedMCTS(mcts.TreePools(B, sims, n_leaves=4), mcts.FusedNetEvaluator(model), use_graph=False, one_launch=one)
            res.append([x.clone() for x in s.search(me, opp, sims)])
        assert all(torch.equal(a, b) for a, b in zip(*res)), B
    print("search_fused_kernel ok")

if stage in ("all", "mlp"):
    model = net.make_net("mlp", seed=3)
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randint(0, 3, (300, 64), device="cuda", generator=g)
    x = torch.stack([(a == 1), (a == 2)], dim=1).reshape(300, 2, 8, 8).to(torch.bfloat16)
    p1 = model.forward_raw(x, fused="pair")
    p2 = model.forward_raw(x, fused="pair2")
    ref = model.forward_raw(x, fused=False)
    torch.cuda.synchronize()
    assert torch.equal(p1, p2)
    assert torch.allclose(p1.float(), ref.float(), atol=3e-2, rtol=3e-2)
    print("mlp_pair / mlp_pair2 ok")

if stage in ("all", "env"):
    me_h, opp_h = po.synthetic_boards(3001, seed=0)
    me, opp = env.to_device_u64(me_h), env.to_device_u64(opp_h)
    mask, act, m2, o2 = env.step_first_legal(me, opp)
    assert np.array_equal(env.to_host_u64(mask), po.legal_mask(me_h, opp_h))
    env.terminal(me, opp)
    env.planes(me, opp)
    print("env kernels ok")

# -DBZ_BOUNDS_CHECK builds export their violation records
import ctypes  # noqa: E402
from betazero_b200 import _lib  # noqa: E402

L = _lib.load()
for name in ("bz_debug_checks_mcts", "bz_debug_checks_selfplay"):
    if hasattr(L, name):
        torch.cuda.synchronize()
        rec = (ctypes.c_int * 4)()
        assert getattr(L, name)(rec) == 0
        print(f"{name}: first failed check {rec[0]} (block {rec[1]}, thread {rec[2]}), violations {rec[3]}")
        assert rec[3] == 0, list(rec)
