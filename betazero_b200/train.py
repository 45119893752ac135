"""Training step on self-play records ("next" rows 1-2 of SURVEY.md section 8f).

Mirrors the reference's supervised loop where it has one (Adam + cross-entropy, batch 128, lr 1e-4:
src/tic_tac_toe/SL/train.py:85-113,182-192; 8-fold dihedral augmentation :27-36; model file
saved at the end :204-214) but with AlphaZero targets: the policy head is fitted to the MCTS visit
distribution pi (soft-label cross-entropy) and the value head to the game result z.
The forward/backward is plain PyTorch (the net is library code); the data path -- bitboards to
planes (K6) and the symmetry augmentation -- runs through the CUDA library.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib, env

N_ACTIONS = 65


def augment(me: torch.Tensor, opp: torch.Tensor, pi: torch.Tensor, sym: torch.Tensor, size: int = 8):
    """Apply dihedral transform ``sym[i]`` (0..7, train.py:27-36) to every (board, pi) record."""
    n = me.numel()
    me_o, opp_o, pi_o = torch.empty_like(me), torch.empty_like(opp), torch.empty_like(pi)
    L = _lib.load()
    _lib.check(L.bz_reversi_symmetry(_lib.dptr(me), _lib.dptr(opp), _lib.dptr(pi.contiguous()), _lib.dptr(sym),
                                     _lib.dptr(me_o), _lib.dptr(opp_o), _lib.dptr(pi_o), n, size, _lib.stream_ptr()),
               "bz_reversi_symmetry")
    return me_o, opp_o, pi_o


REFERENCE_TRANSFORMS = (0, 1, 2, 3, 4, 5, 6, 5)  # the reference's eight lambdas: its 8th, flip(0).t(), IS rot90 x3 (train.py:35)


def expand_with_transforms(me: torch.Tensor, opp: torch.Tensor, pi: torch.Tensor, z: torch.Tensor | None = None, size: int = 8,
                           dedup: bool = True, transforms=tuple(range(8))):
    """The reference's dataset expansion (``TicTacToeDataset.expand_with_transforms``, SL/train.py:23-50) for (board, pi)
    records: every record under the eight transforms, record-major / transform-minor like the reference's two loops, and
    -- ``dedup`` -- only the FIRST occurrence of every distinct (state, action) pair is kept (its ``unique_data`` set).
    ``transforms``: the engine's eight distinct dihedral transforms, or ``REFERENCE_TRANSFORMS`` for the reference's own
    list, whose 8th entry duplicates the 6th and therefore never survives the dedup.
    Returns (me, opp, pi[, z]) of the kept records in that order, and the index of the source record of each."""
    n, T = me.numel(), len(transforms)
    dev = me.device
    src = torch.arange(n, device=dev).repeat_interleave(T)
    sym = torch.tensor(list(transforms), dtype=torch.uint8, device=dev).repeat(n)
    me8, opp8, pi8 = augment(me[src].contiguous(), opp[src].contiguous(), pi[src].contiguous(), sym, size)
    if dedup and n:
        h = torch.empty(n * T, dtype=torch.int64, device=dev)
        L = _lib.load()
        _lib.check(L.bz_record_hash(_lib.dptr(me8), _lib.dptr(opp8), _lib.dptr(pi8), _lib.dptr(h), n * T, _lib.stream_ptr()),
                   "bz_record_hash")
        order = torch.argsort(h, stable=True)  # equal records become neighbours, in their original order
        hs = h[order]
        same_hash = hs[1:] == hs[:-1]
        a, b = order[1:], order[:-1]
        same_rec = same_hash & (me8[a] == me8[b]) & (opp8[a] == opp8[b]) & (pi8[a] == pi8[b]).all(dim=1)
        if bool((same_hash & ~same_rec).any()):
            # two different records share a 64-bit hash (never seen; ~1e-8 for millions of records): equal records may
            # then not be neighbours -- fall back to an exact row-wise unique
            rows = torch.cat([me8[:, None], opp8[:, None], pi8.view(torch.int32).long()], dim=1)
            _, inv = torch.unique(rows, dim=0, return_inverse=True)
            first = torch.full((int(inv.max()) + 1,), n * T, dtype=torch.int64, device=dev).scatter_reduce(
                0, inv, torch.arange(n * T, device=dev), reduce="amin")
            keep = torch.zeros(n * T, dtype=torch.bool, device=dev)
            keep[first] = True
        else:
            keep = torch.ones(n * T, dtype=torch.bool, device=dev)
            keep[a[same_rec]] = False  # a later copy of its predecessor in the run
        me8, opp8, pi8, src = me8[keep], opp8[keep], pi8[keep], src[keep]
    out = (me8, opp8, pi8) + ((z[src],) if z is not None else ())
    return out + (src,)


def make_batch(replay: dict, idx: torch.Tensor, size: int = 8, augment_seed: int | None = None):
    """(planes bf16 [b,2,8,8], pi f32 [b,65], z f32 [b]) for the records ``idx`` of a replay dict."""
    me, opp = replay["me"][idx].contiguous(), replay["opp"][idx].contiguous()
    pi = replay["pi"][idx].contiguous()
    if augment_seed is not None:
        g = torch.Generator(device=me.device).manual_seed(int(augment_seed))
        sym = torch.randint(0, 8, (me.numel(),), dtype=torch.uint8, device=me.device, generator=g)
        me, opp, pi = augment(me, opp, pi, sym, size)
    return env.planes(me, opp), pi, replay["z"][idx].float()


def loss_fn(net: torch.nn.Module, planes: torch.Tensor, pi: torch.Tensor, z: torch.Tensor):
    logits, value = net(planes)
    policy_loss = -(pi * F.log_softmax(logits.float(), dim=-1)).sum(-1).mean()
    value_loss = F.mse_loss(value.float(), z)
    return policy_loss + value_loss, policy_loss.detach(), value_loss.detach()


def train_step(net: torch.nn.Module, opt: torch.optim.Optimizer, planes, pi, z) -> dict:
    net.train()
    opt.zero_grad(set_to_none=True)
    loss, pl, vl = loss_fn(net, planes, pi, z)
    loss.backward()
    opt.step()
    net.eval()
    return {"loss": float(loss.detach()), "policy_loss": float(pl), "value_loss": float(vl)}


def _arch_of(net: torch.nn.Module) -> dict:
    """what load_net() needs to rebuild the module of a checkpoint"""
    from . import net as netmod

    if isinstance(net, netmod.PolicyValueMLP):
        return {"kind": "mlp", "in_features": net.fc1.in_features, "hidden": net.fc1.out_features, "n_actions": net.n_actions}
    if isinstance(net, netmod.PolicyValueResNet):
        return {"kind": "resnet", "channels": net.stem[0].out_channels, "blocks": len(net.tower), "n_actions": net.p_fc.out_features}
    return {"kind": type(net).__name__}


def save_checkpoint(path: str, net: torch.nn.Module, opt: torch.optim.Optimizer | None = None, iteration: int = 0,
                    extra: dict | None = None) -> None:
    """state_dict checkpoint (loads with weights_only=True), unlike the reference's whole-module
    pickle (train.py:204-214) which torch >= 2.6 refuses to load by default.  The architecture is stored beside the
    weights so that a player can be built from the path alone (AIPlayer(path_to_model, symbol), players.py:77-81)."""
    torch.save({"model": net.state_dict(), "optimizer": opt.state_dict() if opt is not None else None,
                "iteration": int(iteration), "arch": _arch_of(net), "extra": extra or {}}, path)


def load_checkpoint(path: str, net: torch.nn.Module, opt: torch.optim.Optimizer | None = None) -> int:
    ck = torch.load(path, map_location="cpu", weights_only=True)
    net.load_state_dict(ck["model"])
    if opt is not None and ck.get("optimizer") is not None:
        opt.load_state_dict(ck["optimizer"])
    if hasattr(net, "prepare_inference") and getattr(net, "_head", None) is not None:
        net.prepare_inference()  # refresh the fused inference buffers in place (captured graphs keep working)
    return int(ck.get("iteration", 0))


def load_net(path: str, device="cuda", dtype=torch.bfloat16) -> torch.nn.Module:
    """Rebuild the module a checkpoint was saved from (its ``arch`` record) and load the weights: the equivalent of the
    reference's ``torch.load(path_to_model)`` (players.py:80) for state_dict checkpoints."""
    from . import net as netmod

    ck = torch.load(path, map_location="cpu", weights_only=True)
    arch = ck.get("arch") or {}
    if arch.get("kind") == "mlp":
        m = netmod.PolicyValueMLP(arch["in_features"], arch["hidden"], arch["n_actions"])
    elif arch.get("kind") == "resnet":
        m = netmod.PolicyValueResNet(arch["channels"], arch["blocks"], arch["n_actions"])
    else:
        raise ValueError(f"checkpoint {path!r} does not describe a known architecture: {arch!r}")
    m.load_state_dict(ck["model"])
    return m.to(device=device, dtype=dtype).eval()
