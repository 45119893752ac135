"""One 800-simulation one-launch search of 4096 Reversi trees (for ncu captures of search_fused_kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import _lib, env, mcts, net as netmod
B, S = int(os.environ.get("GAMES", "4096")), int(os.environ.get("SIMS", "800"))
model = netmod.make_net("mlp", seed=0)
if os.environ.get("LOGIT_SCALE"):  # sharpened priors: deeper trees (bench.py depth_sweep)
    with torch.no_grad():
        model.policy.weight.mul_(float(os.environ["LOGIT_SCALE"]))
        model.policy.bias.mul_(float(os.environ["LOGIT_SCALE"]))
me, opp, _ = env.reversi_init(B)
L = _lib.load()
s = mcts.BatchedMCTS(mcts.TreePools(B, S, n_leaves=4), mcts.FusedNetEvaluator(model), use_graph=False)
for _ in range(int(os.environ.get("REPS", "2"))):
    s.reset(me, opp)
    _lib.check(L.bz_mcts_search_fused(s.pools._ref, _lib.dptr(model._image_pair), _lib.dptr(s.prior_w), S // 4, _lib.stream_ptr()), "fused")
torch.cuda.synchronize()
s.check_errors()
print("ok", s.stats())
