#!/usr/bin/env python
"""GroupNorm launches only (for ncu): SD1.5 level-0/1 shapes at CFG batch 16."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gyre_b200 import _native as N
dev = torch.device("cuda", 0)
N.load()
for (HW, C) in ((4096, 320), (4096, 960), (1024, 640)):
    x = torch.randn(16, HW, C, device=dev).half()
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    for _ in range(2):
        N.groupnorm(x, g, b, 32, 1e-5, True)
torch.cuda.synchronize()
