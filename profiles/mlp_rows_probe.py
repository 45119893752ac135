"""Forward time of the MLP variants vs batch rows (graph of 16 forwards, back to back, no PDL overlap)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import net as netmod
m = netmod.make_net("mlp", seed=0)
for rows in (4096, 8192, 16384, 32768, 65536, 131072):
    x = (torch.rand((rows, 2, 8, 8), device="cuda") > 0.6).to(torch.bfloat16)
    out = torch.empty((rows, 72), dtype=torch.bfloat16, device="cuda")
    res = []
    for mode in ("pair", "pair2", False):
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
        res.append(f"{mode}: {e0.elapsed_time(e1) / 320 * 1e3:6.2f} us")
    print(f"rows {rows:6d}  " + "   ".join(res))
