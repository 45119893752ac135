"""Fine-grained timeline of ONE tree phase (post-net backup + the four descents) of search_fused_kernel: warp 0 of CTA 0,
iteration 150 (library built with -DBZ_TREE_TRACE=0, profiles/build_variant.sh).  Tags: 70 rows published, 71 backup done,
5 root edges requested, 6 root choices made, 10 + d level-d operands in use, 30 + d level-d argmax resolved, 50 leaf moves
applied, 72 descents done."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import _lib, env, mcts, net as netmod
B, S = int(os.environ.get("GAMES", "4096")), 800
model = netmod.make_net("mlp", seed=0)
if os.environ.get("LOGIT_SCALE"):
    with torch.no_grad():
        model.policy.weight.mul_(float(os.environ["LOGIT_SCALE"]))
        model.policy.bias.mul_(float(os.environ["LOGIT_SCALE"]))
me, opp, _ = env.reversi_init(B)
L = _lib.load()
s = mcts.BatchedMCTS(mcts.TreePools(B, S, n_leaves=4), mcts.FusedNetEvaluator(model), use_graph=False)
s.reset(me, opp)
_lib.check(L.bz_mcts_search_fused(s.pools._ref, _lib.dptr(model._image_pair), None, S // 4, _lib.stream_ptr()), "fused")
torch.cuda.synchronize()
t = np.zeros(64, np.int64)
n = ctypes.c_int(0)
L.bz_tree_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert L.bz_tree_debug_trace(t.ctypes.data, ctypes.byref(n)) == 0
ev = t.reshape(32, 2)
t0 = None
for tag, clk in ev:
    if clk == 0:
        continue
    t0 = clk if t0 is None else t0
    print(f"tag {int(tag):3d}  {(clk - t0) / 1965.0:7.2f} us")
