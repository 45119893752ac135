import sys, time
sys.path.insert(0, '/root/repo')
import torch
from betazero_b200 import mcts, net, selfplay
model = net.make_net("mlp", seed=0)
for reuse, noise in ((True, 0.3), (True, 0.0), (False, 0.3)):
    sp = selfplay.BatchedSelfPlay(4096, 800, mcts.FusedNetEvaluator(model), temp_plies=8, seed=7, dirichlet_alpha=noise, n_leaves=4, reuse=reuse)
    sp.prepare()
    t = time.time()
    for i in range(300):
        sp.play_move()
        if i % 50 == 49:
            sp.mcts.check_errors()
    torch.cuda.synchronize()
    st = sp.stats()
    print("reuse", reuse, "noise", noise, "300 plies in", round(time.time() - t, 2), "s", st["games"], "games", st["dropped"], "dropped",
          "max arena units", int(sp.pools.arena_used.max()), "of", sp.pools.arena_units, "max inherited", int(sp.pools.inherited.max()))
