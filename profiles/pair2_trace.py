"""Timeline of the leader CTA of pair 0 of the two-tile pair MLP kernel (needs a library built with -DBZ_MLP_TRACE)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import _lib, net as netmod
m = netmod.make_net("mlp", seed=0)
x = (torch.rand((16384, 2, 8, 8), device="cuda") > 0.6).to(torch.bfloat16)
for _ in range(5):
    m.forward_raw(x, fused="pair2")
torch.cuda.synchronize()
L = _lib.load()
buf = (ctypes.c_longlong * 96)()
L.bz_mlp_pair2_debug_trace.argtypes = [ctypes.c_void_p]
L.bz_mlp_pair2_debug_trace(buf)
names = {0: "control: setup done", 60: "control: loop done"}
for l in range(4):
    for t in range(2):
        names[1 + 8 * l + 4 * t] = f"control: L{l} tile {t} local operands + weights ready"
        names[2 + 8 * l + 4 * t] = f"control: L{l} tile {t} peer ready"
        names[3 + 8 * l + 4 * t] = f"control: L{l} tile {t} MMAs issued + commit"
        names[40 + 4 * l + 2 * t] = f"   epilogue warp 0: L{l} tile {t} accumulators complete"
        names[41 + 4 * l + 2 * t] = f"   epilogue warp 0: L{l} tile {t} operand stored + arrived"
t0 = buf[0]
for t, n in sorted((buf[i] - t0, names[i]) for i in names if buf[i]):
    print(f"{t:7d} clk  {t / 1.965e3:6.2f} us  {n}")
