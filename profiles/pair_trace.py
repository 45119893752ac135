"""Timeline of CTA 0 (leader of pair 0) of the CTA-pair MLP kernel (needs a library built with -DBZ_MLP_TRACE)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import _lib, net as netmod
m = netmod.make_net("mlp", seed=0)
x = (torch.rand((4096, 2, 8, 8), device="cuda") > 0.6).to(torch.bfloat16)
for _ in range(5):
    m.forward_raw(x, fused="pair")
torch.cuda.synchronize()
L = _lib.load()
buf = (ctypes.c_longlong * 64)()
L.bz_mlp_pair_debug_trace.argtypes = [ctypes.c_void_p]
L.bz_mlp_pair_debug_trace(buf)
names = {0: "entry", 1: "previous kernel complete (PDL wait returned)", 2: "cluster sync + x landed", 23: "last epilogue done", 24: "exit"}
for l in range(4):
    names[3 + 5 * l] = f"L{l} operands fenced (CTA barrier)"
    names[4 + 5 * l] = f"L{l} weights landed"
    names[5 + 5 * l] = f"L{l} both CTAs ready"
    names[6 + 5 * l] = f"L{l} MMAs issued + commit"
    names[7 + 5 * l] = f"L{l} accumulators complete"
names[25] = "x copies issued"; names[26] = "cluster sync done"
for l in range(4):
    names[3 + 5 * l] = f"L{l} warp 0: weights landed, arriving"
    for w, o in ((0, 0), (15, 16)):
        if l < 3:
            names[30 + 4 * l + o] = f"L{l} warp {w}: accumulators in registers"
        names[31 + 4 * l + o] = f"L{l} warp {w}: A operand part stored" if l else f"L{l} warp {w}: x copies landed"
        names[32 + 4 * l + o] = f"L{l} warp {w}: fenced"
t0 = buf[0]
for t, n in sorted((buf[i] - t0, names[i]) for i in names if buf[i]):
    print(f"{t:7d} clk  {t / 1.965e3:6.2f} us  {n}")
# kernel alone, back to back
import time
out = torch.empty((4096, 72), dtype=torch.bfloat16, device="cuda")
for mode in (True, "pair"):
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        m.forward_raw(x, out=out, fused=mode)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(16):
                m.forward_raw(x, out=out, fused=mode)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"mode {mode}: {e0.elapsed_time(e1) / 320 * 1e3:.2f} us per forward (graph of 16, back to back)")
