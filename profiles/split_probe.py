"""Time split of one MCTS iteration at the headline config: tree kernel only vs net only vs both
(each replayed from a CUDA graph of 16 iterations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import env, mcts, net as netmod

B, S = 4096, 800
model = netmod.make_net("mlp", seed=0)
me, opp, _ = env.reversi_init(B)

class Static:
    prior_mode = mcts.PRIOR_LOGITS_BF16
    stride = 72
    def bind(self, pools):
        self.out = torch.randn((B, 72), device="cuda").to(torch.bfloat16)
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

for name, ev in (("tree only (static logits)", Static()), ("tree + net", mcts.FusedNetEvaluator(model))):
    pools = mcts.TreePools(B, S)
    s = mcts.BatchedMCTS(pools, ev, graph_unroll=16)
    s.prepare()
    def one():
        s.reset(me, opp)
        s.run(S)
    ms = timed(one)
    print(f"{name}: {ms:.2f} ms per 800-iteration search = {ms / S * 1e3:.2f} us/iteration")

# net only, graph of 16 forwards
x = torch.zeros((B, 2, 8, 8), dtype=torch.bfloat16, device="cuda")
out = torch.zeros((B, 72), dtype=torch.bfloat16, device="cuda")
model.prepare_inference()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(3):
        model.forward_raw(x, out=out)
torch.cuda.current_stream().wait_stream(st)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(16):
        model.forward_raw(x, out=out)
ms = timed(lambda: [g.replay() for _ in range(50)])
print(f"net only: {ms / 800 * 1e3:.2f} us/forward (4 GEMM launches)")
