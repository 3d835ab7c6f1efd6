import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gyre_b200 import _native as N
dev = torch.device("cuda", 0); torch.cuda.set_device(dev); N.load()
M, Nn, K = 65536, 320, 320
a = torch.randn(M, K, device=dev).half(); w = torch.randn(Nn, K, device=dev).half() / math.sqrt(K)
bias = torch.randn(Nn, device=dev); r = torch.randn(M, Nn, device=dev).half()
for i in range(3): N.gemm(a, w, bias=bias, residual=r)
x = torch.randn(16, 64, 64, 320, device=dev).half()
wp = N.pack_conv3x3(torch.randn(320, 320, 3, 3, device=dev).half() / 50)
for i in range(3): N.conv3x3(x, wp, 320, bias=bias)
torch.cuda.synchronize()
