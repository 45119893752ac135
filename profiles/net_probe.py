"""Runs the policy/value net forward (library GEMM path and the fused tcgen05 kernel) on 4096 leaves; used under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import net as netmod
m = netmod.make_net("mlp", seed=0)
x = (torch.rand((4096, 2, 8, 8), device="cuda") > 0.6).to(torch.bfloat16)
out = torch.zeros((4096, 72), dtype=torch.bfloat16, device="cuda")
for _ in range(5):
    m.forward_raw(x, out=out, fused=False)
    m.forward_raw(x, out=out, fused=True)
torch.cuda.synchronize()
print("ok", float(out.float().abs().sum()))
