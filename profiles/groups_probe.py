"""Does splitting the 4096 games of a GPU into independent groups on separate streams help?
Each group has its own pools / CUDA graph; groups are replayed round-robin."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import mcts, net as netmod, selfplay

TOTAL, S = 4096, 800
model = netmod.make_net("mlp", seed=0)
for ng in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(ng)]
    sps = []
    for g in range(ng):
        with torch.cuda.stream(streams[g]):
            sp = selfplay.BatchedSelfPlay(TOTAL // ng, S, mcts.FusedNetEvaluator(model), temp_plies=8, seed=1, rank=g, world=ng)
            sp.prepare()
            sps.append(sp)
    torch.cuda.synchronize()

    def move():
        # reset + first select per group, then interleave graph replays, then the tail
        for g, sp in enumerate(sps):
            with torch.cuda.stream(streams[g]):
                sp.mcts.reset(sp.me, sp.opp)
                sp.mcts.select()
        inner = S - 1
        m = sps[0].mcts
        for _ in range(inner // m.unroll):
            for g, sp in enumerate(sps):
                with torch.cuda.stream(streams[g]):
                    sp.mcts._graph.replay()
        for g, sp in enumerate(sps):
            with torch.cuda.stream(streams[g]):
                for _ in range(inner % m.unroll):
                    sp.mcts.evaluate(); sp.mcts.step()
                sp.mcts.evaluate(); sp.mcts.expand_backup()
                sp.advance()

    for _ in range(3):
        move()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 6
    for _ in range(n):
        move()
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"groups {ng}: {ms:.2f} ms per ply of {TOTAL} games = {TOTAL * S / ms / 1e3:.1f} M sims/s")
    del sps
