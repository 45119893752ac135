"""Lanes per tree (32 / 16 / 8) at the headline batch: full search (tree kernel + pair MLP kernel, PDL) and tree
kernel alone (static logits), per MCTS iteration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import env, mcts, net as netmod

B, S = int(os.environ.get("GAMES", "4096")), 800
model = netmod.make_net("mlp", seed=0)
me, opp, _ = env.reversi_init(B)

class Static:
    prior_mode = mcts.PRIOR_LOGITS_BF16
    stride = 72
    def bind(self, pools):
        g = torch.Generator(device="cuda").manual_seed(3)
        self.out = torch.randn((B, 72), device="cuda", generator=g).to(torch.bfloat16)
        self.value = torch.zeros(1, device="cuda")
        return self.out, self.value
    def __call__(self, pools):
        pass

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

ref = {}
for G in (32, 16, 8):
    for name, ev in (("tree only", Static()), ("tree + net", mcts.FusedNetEvaluator(model))):
        pools = mcts.TreePools(B, S, group_lanes=G, prior_mode=mcts.PRIOR_LOGITS_BF16, eval_stride=72)
        s = mcts.BatchedMCTS(pools, ev, graph_unroll=16)
        s.prepare()
        def one():
            s.reset(me, opp)
            s.run(S)
        ms = timed(one)
        cnt = s.root_policy()[0].clone()
        s.check_errors()
        same = ""
        if name in ref:
            same = f", visit counts identical to G=32: {bool(torch.equal(ref[name], cnt))}"
        else:
            ref[name] = cnt
        print(f"G={G:2d} {name}: {ms:.2f} ms per search = {ms / S * 1e3:.2f} us/iteration, mean depth {s.stats()['mean_depth']:.2f}{same}")
