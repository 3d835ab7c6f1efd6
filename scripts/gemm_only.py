#!/usr/bin/env python
"""One GEMM launch for ncu: python scripts/gemm_only.py M N K [res]"""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gyre_b200 import _native as N
dev = torch.device("cuda", 0)
N.load()
M, Nn, K = (int(v) for v in sys.argv[1:4])
res = len(sys.argv) > 4
a = torch.randn(M, K, device=dev).half()
w = torch.randn(Nn, K, device=dev).half() * (1 / math.sqrt(K))
bias = torch.randn(Nn, device=dev)
r = torch.randn(M, Nn, device=dev).half() if res else None
for _ in range(3):
    N.gemm(a, w, bias=bias, residual=r)
torch.cuda.synchronize()
