"""CPU, world_size 2, gloo: the host-side multi-GPU logic (game sharding, weight broadcast,
replay gather).  The GPU box runs the same code over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from betazero_b200 import dist as bzd
    from betazero_b200 import net

    r, w, _ = bzd.init("gloo")
    assert (r, w) == (rank, world) and bzd.world_size() == world
    # 1. weight broadcast: rank 1's net becomes rank 0's
    model = net.make_net("mlp", seed=rank, device="cpu", dtype=torch.float32)
    ref = net.make_net("mlp", seed=0, device="cpu", dtype=torch.float32)
    nbytes = bzd.broadcast_weights(model, src=0)
    same = all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), ref.state_dict().values()))
    # 2. replay gather: variable-length shards, ordered by (game, ply) independent of rank layout
    slots = 3
    ids = bzd.game_ids(rank, world, slots)
    n = 4 + rank  # ragged
    shard = {
        "game": ids.repeat_interleave(2)[:n].clone(),
        "ply": torch.arange(n, dtype=torch.int16) % 2,
        "me": torch.arange(n, dtype=torch.int64) + 100 * rank,
        "pi": torch.full((n, 65), float(rank)),
        "z": torch.full((n,), rank * 2 - 1, dtype=torch.int8),
    }
    allr = bzd.gather_replay(shard)
    ret[rank] = {"same": same, "nbytes": nbytes, "ids": ids.tolist(),
                 "game": allr["game"].tolist(), "ply": allr["ply"].tolist(), "me": allr["me"].tolist(),
                 "z": allr["z"].tolist(), "pi0": allr["pi"][:, 0].tolist(), "dtypes": {k: str(v.dtype) for k, v in allr.items()}}
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    assert r0["same"] and r1["same"] and r0["nbytes"] == r1["nbytes"] > 0
    assert r0["ids"] == [0, 1, 2] and r1["ids"] == [3, 4, 5]  # disjoint shards of the global id space
    for k in ("game", "ply", "me", "z", "pi0"):
        assert r0[k] == r1[k]  # every rank sees the same gathered replay
    assert len(r0["game"]) == 4 + 5
    keys = [g * 1024 + p for g, p in zip(r0["game"], r0["ply"])]
    assert keys == sorted(keys)
    assert r0["dtypes"]["z"] == "torch.int8" and r0["dtypes"]["ply"] == "torch.int16"
    # records kept their payloads: rank-0 games carry z = -1 / pi = 0, rank-1 games z = +1 / pi = 1
    for g, z, p in zip(r0["game"], r0["z"], r0["pi0"]):
        assert (z, p) == ((-1, 0.0) if g < 3 else (1, 1.0))


def test_single_process_paths():
    from betazero_b200 import dist as bzd
    from betazero_b200 import net

    assert bzd.world_size() == 1
    assert bzd.broadcast_weights(net.make_net("mlp", device="cpu", dtype=torch.float32)) == 0
    shard = {"game": torch.tensor([2, 1, 1]), "ply": torch.tensor([0, 1, 0], dtype=torch.int16), "z": torch.tensor([1, 0, -1])}
    out = bzd.gather_replay(shard)
    assert out["game"].tolist() == [1, 1, 2] and out["ply"].tolist() == [0, 1, 0] and out["z"].tolist() == [-1, 0, 1]
    assert bzd.game_ids(2, 4, 8, round_=1).tolist() == [48 + i for i in range(8)]


def test_net_fast_path_equals_module_forward():
    from betazero_b200 import net

    for game, shape in (("reversi", (7, 2, 8, 8)), ("ttt", (7, 9))):
        m = net.make_net("mlp", game=game, device="cpu", dtype=torch.float32, seed=4)
        x = torch.randint(-1, 2, shape).float()
        logits, v = m(x)
        raw = m.forward_raw(x)
        A = m.n_actions
        assert raw.shape == (7, m.raw_width) and m.raw_width % 8 == 0
        assert torch.allclose(raw[:, :A], logits, atol=1e-5) and torch.allclose(torch.tanh(raw[:, A]), v, atol=1e-5)
    r = net.make_net("resnet", device="cpu", dtype=torch.float32)
    lg, v = r(torch.zeros(3, 2, 8, 8))
    assert lg.shape == (3, 65) and v.shape == (3,)
    assert net.matmul_flops_per_position(net.make_net("mlp", device="cpu")) == 2 * (128 * 256 + 2 * 256 * 256 + 256 * 66)
