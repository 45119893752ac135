"""One 800-iteration search of 4096 Reversi trees with the default pipeline (for ncu captures: cheaper than bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import env, mcts, net as netmod
B, S = int(os.environ.get("GAMES", "4096")), int(os.environ.get("SIMS", "800"))
model = netmod.make_net("mlp", seed=0)
me, opp, _ = env.reversi_init(B)
s = mcts.BatchedMCTS(mcts.TreePools(B, S, n_leaves=int(os.environ.get("LEAVES", "4"))), mcts.FusedNetEvaluator(model), graph_unroll=16)
s.prepare()
s.reset(me, opp)
s.run(S)
torch.cuda.synchronize()
s.check_errors()
print("ok", s.stats())
