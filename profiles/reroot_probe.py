import torch, sys
sys.path.insert(0, "/root/repo")
from betazero_b200 import mcts, selfplay, net as netmod
net = netmod.make_net("mlp", seed=0)
sp = selfplay.BatchedSelfPlay(4096, 800, mcts.FusedNetEvaluator(net), temp_plies=8, seed=4321, n_leaves=4, reuse=True)
for _ in range(12): sp.play_move()
for i in range(3):
    sp.search(); sp.advance()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(); sp.mcts.advance(sp.last_action, sp.me, sp.opp); r1.record(); torch.cuda.synchronize()
    print("reroot ms", r0.elapsed_time(r1), float(sp.pools.inherited.float().mean()), float(sp.pools.arena_used.float().mean()))
    sp._moves += 1
