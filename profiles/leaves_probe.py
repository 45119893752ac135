"""Leaves per tree and iteration (virtual loss): throughput at the headline config (4096 games, 800 sims/move)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import env, mcts, net as netmod

B, S = int(os.environ.get("GAMES", "4096")), 800
model = netmod.make_net("mlp", seed=0)
me, opp, _ = env.reversi_init(B)

def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for K in [int(x) for x in os.environ.get("LEAVES", "1,2,4,8").split(",")]:
    for G in [int(x) for x in os.environ.get("LANES", "0").split(",")]:
        s = mcts.BatchedMCTS(mcts.TreePools(B, S, n_leaves=K, group_lanes=G), mcts.FusedNetEvaluator(model), graph_unroll=16)
        s.prepare()
        def one():
            s.reset(me, opp)
            s.run(S)
        ms = timed(one)
        cnt = s.root_policy()[0]
        s.check_errors()
        st = s.stats()
        print(f"leaves={K} lanes={G or 'auto'}: {ms:.2f} ms per 800-sim search = {ms / (S // K) * 1e3:.2f} us/iteration, "
              f"{B * S / ms / 1e3:.1f} M sims/s; mean depth {st['mean_depth']:.2f}, root visits {int(cnt[0].sum())}, "
              f"max root share {float((cnt.max(1).values.float() / cnt.sum(1).clamp_min(1)).mean()):.3f}")
        del s
        torch.cuda.empty_cache()
