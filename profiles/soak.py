import sys, time
sys.path.insert(0, '/root/repo')
import torch
from betazero_b200 import mcts, net, selfplay
model = net.make_net("mlp", seed=0)
sp = selfplay.BatchedSelfPlay(4096, 800, mcts.FusedNetEvaluator(model), temp_plies=8, seed=7, dirichlet_alpha=0.3, n_leaves=int(sys.argv[1]) if len(sys.argv) > 1 else 4)
sp.prepare()
t = time.time()
for i in range(200):
    sp.play_move()
    if i % 50 == 49:
        sp.mcts.check_errors()
torch.cuda.synchronize()
print("200 plies in", round(time.time() - t, 2), "s", sp.stats(), sp.mcts.stats())
rp = sp.drain_replay()
print("replay", rp["me"].shape, "z mean", float(rp["z"].float().mean()), "pi rowsum", float(rp["pi"].sum(1).mean()))
