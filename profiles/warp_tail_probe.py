"""Per-warp timing of the LAST step_wave_kernel launch of a search (library built with -DBZ_WARP_PROBE, see
profiles/r2_experiments.sh): when does each tree's warp enter / finish expand+backup / leave the kernel, and how many
level-steps did its descents take.  Answers: how much of the kernel's duration is the tail of its deepest trees?"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import _lib, env, mcts, net as netmod

B, S, K = int(os.environ.get("GAMES", "4096")), 800, int(os.environ.get("LEAVES", "4"))
model = netmod.make_net("mlp", seed=0)
me, opp, _ = env.reversi_init(B)
pools = mcts.TreePools(B, S, n_leaves=K)
s = mcts.BatchedMCTS(pools, mcts.FusedNetEvaluator(model), use_graph=True, graph_unroll=16)
s.prepare()
L = _lib.load()
L.bz_warp_probe.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int]
for upto in (160, 400, 784):  # simulations done before the probed launch: shallow, mid, deep trees
    s.reset(me, opp)
    s.run(upto)  # graph replays (PDL chained) + the trailing eager iterations; the last step kernel is the probed one
    s.select()
    for _ in range(3):
        s.evaluate(); s.step()
    torch.cuda.synchronize()
    n = min(B, 16384)
    t0, t1, t2 = (np.zeros(n, np.uint64) for _ in range(3))
    st = np.zeros(n, np.int32)
    rc = L.bz_warp_probe(t0.ctypes.data, t1.ctypes.data, t2.ctypes.data, st.ctypes.data, n)
    assert rc == 0, rc
    base = t0.min()
    ent, mid, ext = (t0 - base) / 1e3, (t2 - base) / 1e3, (t1 - base) / 1e3
    dur = ext - ent
    q = lambda a: " ".join(f"{np.percentile(a, p):6.2f}" for p in (0, 10, 50, 90, 99, 100))
    print(f"--- after {upto} sims: level-steps per warp  min/p10/p50/p90/p99/max: {q(st)}   mean {st.mean():.2f}")
    print(f"    warp entry  (us after first entry)   {q(ent)}")
    print(f"    expand+backup done (us after own entry) {q(mid - ent)}")
    print(f"    warp exit   (us after first entry)   {q(ext)}")
    print(f"    warp duration (us)                   {q(dur)}   mean {dur.mean():.2f}")
    print(f"    kernel span {ext.max():.2f} us; warps still running at 50/75/90/100% of span: "
          + " ".join(str(int((ext > f * ext.max()).sum())) for f in (0.5, 0.75, 0.9, 0.999)))
    per_step = (ext - mid) / np.maximum(st, 1)
    print(f"    select phase us per level-step       {q(per_step)}")
