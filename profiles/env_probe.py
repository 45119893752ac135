"""Runs the env kernels on 2^26 synthetic boards (working set > L2) a few times; used under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import env

n = 1 << 26
g = torch.Generator(device="cuda").manual_seed(0)
a, b, c = (torch.randint(-(2 ** 63), 2 ** 63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g) for _ in range(3))
me, opp = a & b, ~a & c
del a, b, c
mask = torch.empty_like(me)
outs = (torch.empty_like(me), torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty_like(me), torch.empty_like(me))
act = torch.zeros(n, dtype=torch.uint8, device="cuda")
for _ in range(4):
    env.legal_mask(me, opp, out=mask)
    env.step_first_legal(me, opp, out=outs)
    env.terminal(me, opp)
    env.apply(me, opp, outs[1])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("legal_mask", lambda: env.legal_mask(me, opp, out=mask)), ("step_first_legal", lambda: env.step_first_legal(me, opp, out=outs))):
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(name, f"{ms:.4f} ms  {n / ms / 1e6:.1f} G boards/s")
