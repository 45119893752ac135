"""One full AlphaZero iteration (BASELINE configs[4]): lockstep self-play on every GPU -> replay
gather (NCCL all-gather) -> training steps on the gathered records -> weight broadcast.

    torchrun --nproc-per-node 8 -m betazero_b200.loop --games 4096 --sims 800 --plies 70 --train-steps 8

Self-play is the hot path (hand-written kernels, no collective); the two collectives run once per
iteration (`dist.gather_replay`, `dist.broadcast_weights`).  Every rank trains on the same gathered
batch order with the same seed, so the replicas stay bit-identical and the broadcast is a
consistency guarantee rather than a necessity; with ``--train-on-rank0`` only rank 0 trains.
Prints one JSON line per iteration on rank 0.
"""
from __future__ import annotations

import argparse
import json
import time

import torch

from . import dist as bzd
from . import mcts, net, selfplay, train


def run(args) -> list[dict]:
    rank, world, local = bzd.init()
    if not torch.cuda.is_available():
        raise SystemExit("betazero_b200.loop needs CUDA devices: there is no CPU fallback")
    torch.cuda.set_device(local)
    model = net.make_net(args.net, hidden=args.hidden, seed=args.seed)
    master = net.make_net(args.net, hidden=args.hidden, seed=args.seed, dtype=torch.float32)  # fp32 master weights
    opt = torch.optim.Adam(master.parameters(), lr=args.lr)  # the reference's optimiser family (SL/train.py:87)
    evaluator = mcts.FusedNetEvaluator(model) if hasattr(model, "forward_raw") else mcts.NetEvaluator(model)
    sp = selfplay.BatchedSelfPlay(args.games, args.sims, evaluator, board_size=args.size, temp_plies=args.temp_plies,
                                  seed=args.seed, rank=rank, world=world, n_leaves=args.leaves,
                                  graph_unroll=min(16, max(1, args.sims // args.leaves - 1)))
    sp.prepare()
    out = []
    for it in range(args.iterations):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev[0].record()
        for _ in range(args.plies):
            sp.play_move()
        ev[1].record()
        shard = sp.drain_replay()
        replay = bzd.gather_replay(shard)
        ev[2].record()
        n = replay["me"].shape[0]
        losses = []
        if n and (not args.train_on_rank0 or rank == 0):
            g = torch.Generator(device="cuda").manual_seed(args.seed * 1000 + it)
            for k in range(args.train_steps):
                idx = torch.randint(0, n, (min(args.batch, n),), device="cuda", generator=g)
                planes, pi, z = train.make_batch(replay, idx, size=args.size, augment_seed=args.seed + 31 * it + k)
                losses.append(train.train_step(master, opt, planes, pi, z)["loss"])
            with torch.no_grad():
                for dst, src in zip(model.parameters(), master.parameters()):
                    dst.copy_(src)
        ev[3].record()
        nbytes = bzd.broadcast_weights(model, src=0)
        evaluator.refresh() if hasattr(evaluator, "refresh") else None
        ev[4].record()
        torch.cuda.synchronize()
        sp.mcts.check_errors()
        st = sp.stats()
        line = {
            "iteration": it, "world": world, "games_per_gpu": args.games, "sims_per_move": args.sims, "plies": args.plies,
            "wall_s": time.perf_counter() - t0,
            "ms": {"selfplay": ev[0].elapsed_time(ev[1]), "replay_gather": ev[1].elapsed_time(ev[2]),
                   "train": ev[2].elapsed_time(ev[3]), "weight_broadcast": ev[3].elapsed_time(ev[4])},
            "sims_per_sec_per_gpu": args.games * args.sims * args.plies / (ev[0].elapsed_time(ev[1]) * 1e-3),
            "replay_records_gathered": int(n), "local_records": int(shard["me"].shape[0]),
            "games_finished_local": st["games"], "loss_first_last": [losses[0], losses[-1]] if losses else None,
            "broadcast_bytes": nbytes,
        }
        out.append(line)
        if rank == 0:
            print(json.dumps(line), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=4096)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--leaves", type=int, default=4, help="virtual-loss descents per tree and iteration (1 = sequential search)")
    ap.add_argument("--plies", type=int, default=70, help="lockstep plies per iteration (a game lasts ~60)")
    ap.add_argument("--size", type=int, default=8)
    ap.add_argument("--iterations", type=int, default=1)
    ap.add_argument("--train-steps", type=int, default=8)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--lr", type=float, default=1e-4)  # SL/train.py:190
    ap.add_argument("--temp-plies", type=int, default=8)
    ap.add_argument("--net", default="mlp")
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--train-on-rank0", action="store_true")
    run(ap.parse_args())
    if bzd.world_size() > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
