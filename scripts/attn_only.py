#!/usr/bin/env python
"""One self-attention launch at the SD1.5 level-0 shape (for ncu): python scripts/attn_only.py [variant]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gyre_b200 import _native as N
dev = torch.device("cuda", 0)
N.load()
if len(sys.argv) > 1:
    N.set_tunable("ATT_VARIANT", int(sys.argv[1]))
B, heads, Nq, d = 16, 8, 4096, 40
C = heads * d
qkv = torch.randn(B, Nq, 3 * C, device=dev).half()
for _ in range(2):
    N.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads)
torch.cuda.synchronize()
