"""Timeline of one iteration of search_fused_kernel in CTA 0 (library built with -DBZ_FUSED_TRACE=<iteration>, see
profiles/build_variant.sh): island 0 / island 1 phase boundaries and the control warp's MMA issue times, in us."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import _lib, env, mcts, net as netmod
B, S = int(os.environ.get("GAMES", "4096")), 800
model = netmod.make_net("mlp", seed=0)
if os.environ.get("LOGIT_SCALE"):  # sharpened priors: deeper trees (bench.py depth_sweep)
    with torch.no_grad():
        model.policy.weight.mul_(float(os.environ["LOGIT_SCALE"]))
        model.policy.bias.mul_(float(os.environ["LOGIT_SCALE"]))
me, opp, _ = env.reversi_init(B)
if os.environ.get("PLIES"):  # mid-game roots: play lockstep self-play plies first (the bench region is plies 5-24)
    from betazero_b200 import selfplay
    sp = selfplay.BatchedSelfPlay(B, S, mcts.FusedNetEvaluator(model), temp_plies=8, seed=1234, n_leaves=4)
    for _ in range(int(os.environ["PLIES"])):
        sp.play_move()
    me, opp = sp.me.clone(), sp.opp.clone()
    del sp
L = _lib.load()
s = mcts.BatchedMCTS(mcts.TreePools(B, S, n_leaves=4), mcts.FusedNetEvaluator(model), use_graph=False)
for _ in range(2):
    s.reset(me, opp)
    _lib.check(L.bz_mcts_search_fused(s.pools._ref, _lib.dptr(model._image_pair), _lib.dptr(s.prior_w), S // 4, _lib.stream_ptr()), "fused")
torch.cuda.synchronize()
print("stats", s.stats())
t = np.zeros(96, np.int64)
L.bz_fused_debug_trace.argtypes = [ctypes.c_void_p]
assert L.bz_fused_debug_trace(t.ctypes.data) == 0
t = t.reshape(3, 32).astype(np.float64)
t0 = t[t > 0].min()
us = lambda x: (x - t0) / 1965.0
names = ["iter start", "engine free", "layer-0 operand stored", "L0 acc ready", "L0 epilogue done", "L1 acc ready", "L1 epilogue done",
         "L2 acc ready", "L2 epilogue done", "head acc ready", "head rows stored", "island barrier passed", "expand+backup done", "select done"]
idx = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]
for I in range(2):
    print(f"island {I}: " + "; ".join(f"{n} {us(t[I, i]):.2f}" for n, i in zip(names, idx) if t[I, i] > 0))
for I in range(2):
    print(f"control, island {I} job: " + "; ".join(f"L{l}: operands {us(t[2, I * 16 + 3 * l]):.2f} peer {us(t[2, I * 16 + 3 * l + 1]):.2f} issued {us(t[2, I * 16 + 3 * l + 2]):.2f}" for l in range(4)))
