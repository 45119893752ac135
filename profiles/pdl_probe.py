"""Per-iteration time at the headline config: library GEMMs vs the fused tcgen05 MLP kernel, with and
without programmatic dependent launch between the two kernels of an iteration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import _lib, env, mcts, net as netmod

B, S = 4096, 800
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

ref = None
for name, use_kernel, pdl in (("library GEMMs", False, False), ("library GEMMs + PDL attribute on the step kernel", False, True), ("fused MLP kernel", True, False), ("fused MLP kernel + PDL", True, True),
                              ("pair MLP kernel", "pair", False), ("pair MLP kernel + PDL", "pair", True)):
    pools = mcts.TreePools(B, S)
    s = mcts.BatchedMCTS(pools, mcts.FusedNetEvaluator(model, use_kernel=use_kernel, pdl=pdl), graph_unroll=16)
    s.prepare()
    def one():
        s.reset(me, opp)
        s.run(S)
    ms = timed(one)
    cnt = s.root_policy()[0].clone()
    s.check_errors()
    if name == "fused MLP kernel":
        ref = cnt
    elif "MLP kernel" in name:
        print("   PDL result identical to non-PDL:", bool(torch.equal(ref, cnt)))
    print(f"{name}: {ms:.2f} ms per search = {ms / S * 1e3:.2f} us/iteration, visits ok: {bool((cnt.sum(1) == S - 1).all())}")
_lib.set_pdl(False)
