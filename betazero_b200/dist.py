"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU
box, gloo in the CPU tests).

Self-play needs NO data-path collective: games are independent units (the reference has no
cross-game state: each ReversiTerminal / TicTacToeHeadless owns its board, reversi_terminal.py:11-14,
tic_tac_toe.py:7-11), so rank r of W simply owns game ids ``{r*B + s + k*W*B}``.  Collectives
appear only between iterations: ``broadcast_weights`` (new net from the training rank) and
``gather_replay`` (replay shards to every rank / the training rank).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise from the torchrun environment.  Returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the caller's own output
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def game_ids(rank: int, world: int, slots: int, round_: int = 0) -> torch.Tensor:
    """Global game ids owned by ``rank`` in restart round ``round_`` (matches bz_selfplay_init /
    id_stride): slot s plays ``rank*slots + s + round_*world*slots``."""
    return torch.arange(slots, dtype=torch.int64) + rank * slots + round_ * world * slots


@torch.no_grad()
def broadcast_weights(module: torch.nn.Module, src: int = 0) -> int:
    """Broadcast every parameter and buffer of ``module`` from ``src`` as ONE flat buffer per dtype
    (few MB: launch-latency bound, so one collective instead of one per tensor).  Returns bytes."""
    if world_size() == 1:
        return 0
    tensors = [t for t in list(module.parameters()) + list(module.buffers()) if t.numel()]
    total = 0
    by_dtype: dict[torch.dtype, list[torch.Tensor]] = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dt, ts in by_dtype.items():
        flat = torch.cat([t.detach().reshape(-1) for t in ts])
        dist.broadcast(flat, src=src)
        off = 0
        for t in ts:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
        total += flat.numel() * flat.element_size()
    return total


def pack_records(local: dict) -> tuple[torch.Tensor, list]:
    """Interleave a dict of per-record tensors (equal leading length n) into ONE uint8 [n, record_bytes] buffer
    (fields in sorted key order, the record padded to a multiple of 16 bytes) + the layout to undo it."""
    keys = sorted(local)
    n = int(local[keys[0]].shape[0])
    cols, layout, off = [], [], 0
    for k in keys:
        t = local[k].contiguous()
        width = 1
        for d in t.shape[1:]:
            width *= int(d)
        row_bytes = width * t.element_size()
        layout.append((k, t.dtype, tuple(t.shape[1:]), off, row_bytes))
        cols.append(t.reshape(n, width).view(torch.uint8).reshape(n, row_bytes))
        off += row_bytes
    pad = (-off) % 16
    if pad:
        cols.append(torch.zeros((n, pad), dtype=torch.uint8, device=cols[0].device))
    return torch.cat(cols, dim=1).contiguous(), layout


def unpack_records(buf: torch.Tensor, layout: list) -> dict:
    n = int(buf.shape[0])
    out = {}
    for k, dt, shape, off, row_bytes in layout:
        out[k] = buf[:, off:off + row_bytes].contiguous().view(dt).reshape((n,) + shape)
    return out


last_gather = {}  # bytes / record size of the most recent gather_replay (for the iteration report)


def gather_replay(local: dict) -> dict:
    """All-gather variable-length replay shards (dict of tensors with equal leading length) to
    every rank, ordered by (game id, ply) so the result does not depend on the rank layout.

    Two collectives in all: the shard lengths, then ONE all-gather of the records packed into a single byte buffer
    (288 B per self-play record: me, opp, pi[65], z, game id, ply), padded to the longest shard; one host sync."""
    keys = sorted(local)
    n_local = int(local[keys[0]].shape[0])
    last_gather.clear()
    if world_size() == 1:
        out = {k: local[k] for k in keys}
    else:
        W = world_size()
        dev = local[keys[0]].device
        packed, layout = pack_records(local)
        rec = int(packed.shape[1])
        counts = torch.zeros(W, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, torch.tensor([n_local], dtype=torch.int64, device=dev))
        cnt = counts.cpu().tolist()  # the one host sync: the padded length must be known to size the buffer
        n_max = max(cnt)
        send = packed
        if n_local < n_max:
            send = torch.zeros((n_max, rec), dtype=torch.uint8, device=dev)
            send[:n_local] = packed
        buf = torch.empty((W * n_max, rec), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(buf, send)
        rows = torch.cat([buf[r * n_max: r * n_max + cnt[r]] for r in range(W)]) if n_max else buf
        out = unpack_records(rows, layout)
        last_gather.update(record_bytes=rec, records=sum(cnt), wire_bytes_per_rank=n_max * rec,
                           gathered_bytes=W * n_max * rec)
    if "game" in out and "ply" in out and out["game"].numel():
        order = torch.argsort(out["game"] * 1024 + out["ply"].to(torch.int64))
        out = {k: v[order] for k, v in out.items()}
    return out
