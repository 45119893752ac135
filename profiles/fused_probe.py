"""The one-launch search (bz_mcts_search_fused) against the per-iteration kernels: identical trees, time per search."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import _lib, env, mcts, net as netmod

B, S = int(os.environ.get("GAMES", "4096")), int(os.environ.get("SIMS", "800"))
model = netmod.make_net("mlp", seed=0)
if os.environ.get("LOGIT_SCALE"):  # sharpened priors: deeper trees (bench.py depth_sweep)
    with torch.no_grad():
        model.policy.weight.mul_(float(os.environ["LOGIT_SCALE"]))
        model.policy.bias.mul_(float(os.environ["LOGIT_SCALE"]))
me, opp, _ = env.reversi_init(B)
L = _lib.load()

def searcher():
    s = mcts.BatchedMCTS(mcts.TreePools(B, S, n_leaves=4), mcts.FusedNetEvaluator(model), graph_unroll=16, one_launch=False)
    s.prepare()
    return s

ref = searcher()
ref.reset(me, opp); ref.run(S); torch.cuda.synchronize(); ref.check_errors()
Nr, Wr, Pr = [x.clone() for x in ref.root_edges()]

f = searcher()
def fused_once():
    f.reset(me, opp)
    _lib.check(L.bz_mcts_search_fused(f.pools._ref, _lib.dptr(model._image_pair), _lib.dptr(f.prior_w), S // 4, _lib.stream_ptr()), "fused")
fused_once(); torch.cuda.synchronize(); f.check_errors()
Nf, Wf, Pf = f.root_edges()
print("root N equal:", bool((Nr == Nf).all()), " W equal:", bool((Wr == Wf).all()), " P equal:", bool((Pr == Pf).all()),
      " visits", int(Nf[0].sum()), int(Nr[0].sum()))
print("arena_used equal:", bool((ref.pools.arena_used == f.pools.arena_used).all()), " arenas equal:",
      bool((ref.pools.arena[: 1 << 22] == f.pools.arena[: 1 << 22]).all()))

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def ref_once():
    ref.reset(me, opp); ref.run(S)
for name, fn in (("per-iteration kernels", ref_once), ("one launch", fused_once)):
    ms = timed(fn)
    print(f"{name}: {ms:.3f} ms per {S}-sim search = {ms / (S // 4) * 1e3:.2f} us/iteration, {B * S / ms / 1e3:.1f} M sims/s")
