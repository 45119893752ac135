"""INT32 pipe microbenchmarks (ALU-only vs ALU+FMA dual issue)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from betazero_b200 import env
for variant in (0, 1):
    for blocks in (148 * 4, 148 * 8):
        ops = env.int32_microbench(blocks, 256, 2048, variant)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            env.int32_microbench(blocks, 256, 2048, variant)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"variant {variant} blocks {blocks}: {ops / ms / 1e9:.2f} T int-instr lanes/s ({ms:.3f} ms)")
