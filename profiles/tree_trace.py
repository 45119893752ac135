"""Timeline of one warp (one tree) of the fused step kernel at the headline config
(needs a library built with -DBZ_TREE_TRACE=<cta index>)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import _lib, env, mcts, net as netmod
B, S = 4096, 800
model = netmod.make_net("mlp", seed=0)
me, opp, _ = env.reversi_init(B)
K = int(os.environ.get("LEAVES", "4"))
pools = mcts.TreePools(B, S, n_leaves=K)
s = mcts.BatchedMCTS(pools, mcts.FusedNetEvaluator(model), use_graph=False)
s.reset(me, opp)
s.run(600)            # deep-ish trees, eager so that the last step kernel is the traced one
s.select()
s.evaluate(); s.step()
s.evaluate(); s.step()
torch.cuda.synchronize()
L = _lib.load()
buf = (ctypes.c_longlong * 64)(); n = ctypes.c_int()
L.bz_tree_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
L.bz_tree_debug_trace(buf, ctypes.byref(n))
names = {0: "kernel start", 1: "leaf records + root loaded", 2: "evaluator row in, priors done", 3: "expand/backup stores issued",
         4: "node blocks + links stored", 50: "leaf rules done", 60: "end"}
# the level tags are written by lane 0 = slot 0 of the tree's warp: "level d" is slot 0's depth (= the level-step while it is
# still descending); steps after slot 0 has finished are visible as the gap before "leaf rules done"
t0 = buf[1]
for i in range(n.value):
    tag, t = buf[2 * i], buf[2 * i + 1]
    nm = names.get(tag) or (f"level {tag - 10}: loads issued" if tag < 30 else f"level {tag - 30}: argmax resolved")
    print(f"{t - t0:6d} clk {(t - t0) / 1.965e3:5.2f} us  {nm}")
