#!/usr/bin/env python
"""tcgen05.mma issue/latency microbenchmark (M=128, K=16): clocks per instruction vs N, accumulator count, A source."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gyre_b200 import _native as N
dev = torch.device("cuda", 0)
lib = N.load()
out = torch.zeros(2 * 148, dtype=torch.int64, device=dev)
REPS = 1200
for blocks in (148,):
    for a_tmem in (0, 1):
        for n in (16, 32, 48, 64, 96, 128, 192, 256):
            row = []
            for naccs in ((1, 2, 3) if n <= 128 else (1,)):
                N.check(lib.gyre_b200_debug_mma_bench(n, naccs, a_tmem, REPS, blocks, N.ptr(out), N.stream_ptr(dev)), "mma_bench")
                torch.cuda.synchronize()
                o = out[:2 * blocks].view(blocks, 2).float().mean(0)
                row.append(f"accs={naccs}: {o[0].item() / REPS:6.1f} (issue {o[1].item() / REPS:5.1f})")
            print(f"blocks={blocks:3d} A={'tmem' if a_tmem else 'smem'} N={n:3d}  clk/MMA  " + "  ".join(row), flush=True)
