"""Race check by repetition: the one-launch search of the headline configuration, REPS times on the same roots; every run
must reproduce the first one bit for bit (root statistics, allocator state, every live arena word).  A missed ordering in
the kernel's barrier protocol (row buffer, island barriers, accumulator hand-over) would show up as a difference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import env, mcts, net as netmod
B, S, REPS = int(os.environ.get("GAMES", "4096")), int(os.environ.get("SIMS", "800")), int(os.environ.get("REPS", "40"))
K = int(os.environ.get("LEAVES", "4"))
model = netmod.make_net("mlp", seed=0)
me, opp, _ = env.reversi_init(B)
s = mcts.BatchedMCTS(mcts.TreePools(B, S, n_leaves=K), mcts.FusedNetEvaluator(model), use_graph=False, one_launch=True)
ref, bad = None, 0
for r in range(REPS):
    s.pools.arena.zero_()
    s.reset(me, opp)
    s.run(S)
    torch.cuda.synchronize()
    s.check_errors()
    cur = [x.clone() for x in s.root_edges()] + [s.pools.arena_used.clone(), s.pools.arena.clone()]
    if ref is None:
        ref = cur
    else:
        bad += int(not all(torch.equal(a, b) for a, b in zip(ref, cur)))
print(f"leaves={K}: {REPS} runs of {B} trees x {S} sims, {bad} differ from the first")
