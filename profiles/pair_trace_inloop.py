"""Timeline of CTA 0 of the pair MLP kernel INSIDE the search loop (last launch of a graph-replayed search): shows how
much of its prologue programmatic dependent launch hides behind the tree kernel.  Needs -DBZ_MLP_TRACE."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import _lib, env, mcts, net as netmod
B, S = 4096, 400
model = netmod.make_net("mlp", seed=0)
me, opp, _ = env.reversi_init(B)
s = mcts.BatchedMCTS(mcts.TreePools(B, S), mcts.FusedNetEvaluator(model), graph_unroll=16)
s.prepare()
s.reset(me, opp)
s.run(S - 7)   # ends inside graph replays + eager tail
torch.cuda.synchronize()
L = _lib.load()
buf = (ctypes.c_longlong * 64)()
L.bz_mlp_pair_debug_trace.argtypes = [ctypes.c_void_p]
L.bz_mlp_pair_debug_trace(buf)
names = {0: "entry", 1: "previous kernel complete (PDL wait returned)", 25: "x copies issued", 26: "cluster sync done", 2: "x landed (warp 0)",
         23: "last epilogue done", 24: "exit"}
for l in range(4):
    names[3 + 5 * l] = f"L{l} warp 0: CTA ready, weights landed"
    names[5 + 5 * l] = f"L{l} both CTAs ready"
    names[6 + 5 * l] = f"L{l} MMAs issued + commit"
    names[7 + 5 * l] = f"L{l} accumulators complete"
t0 = buf[0]
for t, n in sorted((buf[i] - t0, names[i]) for i in names if buf[i]):
    print(f"{t:7d} clk  {t / 1.965e3:6.2f} us  {n}")
