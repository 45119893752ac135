"""Timeline of CTA 0 of the pipelined MLP kernel (needs a library built with -DBZ_MLP_TRACE)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import _lib, net as netmod
m = netmod.make_net("mlp", seed=0)
x = (torch.rand((4096, 2, 8, 8), device="cuda") > 0.6).to(torch.bfloat16)
for _ in range(5):
    m.forward_raw(x, fused=True)
torch.cuda.synchronize()
L = _lib.load()
buf = (ctypes.c_longlong * 64)()
L.bz_mlp_debug_trace.argtypes = [ctypes.c_void_p]
L.bz_mlp_debug_trace(buf)
t0 = buf[0]
names = {0: "setup done", 10: "x loaded", 30: "all done"}
for L_ in range(4):
    for h in range(2):
        names[1 + L_ * 2 + h] = f"MMA issued L{L_} h{h}"
for L_ in range(3):
    names[11 + L_ * 4] = f"epi sees D L{L_} h0"; names[12 + L_ * 4] = f"epi done  L{L_} cg0"
    names[13 + L_ * 4] = f"epi sees D L{L_} h1"; names[14 + L_ * 4] = f"epi done  L{L_} cg2"
ev = sorted((buf[i] - t0, names[i]) for i in names if buf[i])
for t, n in ev:
    print(f"{t:7d} clk  {t / 1.965e3:6.2f} us  {n}")
