#!/usr/bin/env python
"""Per-shape kernel timings for the SD1.5 UNet at CFG batch 16 (8 images): every distinct GEMM / conv /
attention / norm shape, timed by CUDA-graph replay (no host overhead) with rotating buffers, next to its roofline
(max of FLOPs / measured tensor peak and algorithmic bytes / measured HBM copy bandwidth)."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gyre_b200 import _native as N  # noqa: E402

PEAK_TF, PEAK_GB = 1403.1, 6454.6
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    d = json.load(open(p))
    PEAK_TF, PEAK_GB = d["bf16_tflops_sustained"], d["hbm_gbs"]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
N.load()
B2 = 16
ROT = 4


def timeit(fn, iters=8, replays=4):
    """CUDA-graph replay of `iters` launches over rotating buffers: no host launch overhead in the number."""
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for i in range(ROT):
            fn(i)
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * replays) * 1e3   # us


rows = []


def report(kind, name, us, flops, bytes_):
    ideal = max(flops / (PEAK_TF * 1e12), bytes_ / (PEAK_GB * 1e9)) * 1e6
    rows.append((kind, name, us, flops / us / 1e6 if flops else 0.0, bytes_ / us / 1e3, ideal, ideal / us))
    print(f"{kind:6s} {name:44s} {us:9.1f} us  {flops / us / 1e6 if flops else 0:7.1f} TF/s {bytes_ / us / 1e3:7.0f} GB/s  "
          f"ideal {ideal:7.1f} us  frac {ideal / us:5.2f}", flush=True)


def bench_gemm(M, Nn, K, res=False, geglu=False, count=1, name=""):
    a = [torch.randn(M, K, device=dev).half() for _ in range(ROT)]
    w = torch.randn(Nn, K, device=dev).half() * (1 / math.sqrt(K))
    bias = torch.randn(Nn, device=dev)
    r = [torch.randn(M, Nn, device=dev).half() for _ in range(ROT)] if res else None
    if geglu:
        w, bias = N.pack_geglu(w, bias)
    us = timeit(lambda i: N.gemm(a[i % ROT], w, bias=bias, residual=r[i % ROT] if res else None, act=1 if geglu else 0))
    nout = Nn // 2 if geglu else Nn
    report("gemm", f"{name} M{M} N{Nn} K{K}{' +res' if res else ''}{' geglu' if geglu else ''} x{count}", us,
           2.0 * M * Nn * K, 2.0 * (M * K + Nn * K + M * nout * (2 if res else 1)))
    return us * count


def bench_conv(B, H, W, Cin, Cout, stride=1, count=1, name=""):
    x = [torch.randn(B, H, W, Cin, device=dev).half() for _ in range(ROT)]
    w = torch.randn(Cout, Cin, 3, 3, device=dev).half() * (1 / math.sqrt(9 * Cin))
    wp = N.pack_conv3x3(w)
    bias = torch.randn(Cout, device=dev)
    us = timeit(lambda i: N.conv3x3(x[i % ROT], wp, Cout, bias=bias, stride=stride))
    Ho, Wo = H // stride, W // stride
    report("conv", f"{name} {B}x{H}x{W} {Cin}->{Cout} s{stride} x{count}", us, 2.0 * 9 * Cin * Cout * B * Ho * Wo,
           2.0 * (B * H * W * Cin + 9 * Cin * Cout + B * Ho * Wo * Cout))
    return us * count


def bench_attn(B, heads, Nq, Nk, d, count=1, name=""):
    C = heads * d
    q = [torch.randn(B, Nq, C, device=dev).half() for _ in range(ROT)]
    k = torch.randn(B, Nk, C, device=dev).half()
    v = torch.randn(B, Nk, C, device=dev).half()
    us = timeit(lambda i: N.attention(q[i % ROT], k, v, heads))
    report("attn", f"{name} B{B} h{heads} Nq{Nq} Nk{Nk} d{d} x{count}", us, 4.0 * B * heads * Nq * Nk * d,
           2.0 * B * C * (2 * Nq + 2 * Nk))
    return us * count


def bench_gn(B, HW, C, count=1, silu=True):
    x = [torch.randn(B, HW, C, device=dev).half() for _ in range(ROT)]
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    us = timeit(lambda i: N.groupnorm(x[i % ROT], g, b, 32, 1e-5, silu))
    report("gn", f"B{B} HW{HW} C{C} x{count}", us, 0.0, 2.0 * 2 * B * HW * C)
    return us * count


def bench_ln(rows_, C, count=1):
    x = [torch.randn(rows_, C, device=dev).half() for _ in range(ROT)]
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    us = timeit(lambda i: N.layernorm(x[i % ROT], g, b))
    report("ln", f"rows{rows_} C{C} x{count}", us, 0.0, 2.0 * 2 * rows_ * C)
    return us * count


tot = {"gemm": 0.0, "conv": 0.0, "attn": 0.0, "gn": 0.0, "ln": 0.0}
# transformer blocks: (level HW, C, count): 5 at level 0/1/2, 1 mid
for HW, C, cnt in ((4096, 320, 5), (1024, 640, 5), (256, 1280, 5), (64, 1280, 1)):
    M = B2 * HW
    tot["gemm"] += bench_gemm(M, C, C, count=2 * cnt, name="proj_in/q2")
    tot["gemm"] += bench_gemm(M, 3 * C, C, count=cnt, name="qkv")
    tot["gemm"] += bench_gemm(M, C, C, res=True, count=3 * cnt, name="o1/o2/proj_out")
    tot["gemm"] += bench_gemm(B2 * 77, 2 * C, 768, count=cnt, name="kv2")
    tot["gemm"] += bench_gemm(M, 8 * C, C, geglu=True, count=cnt, name="geglu")
    tot["gemm"] += bench_gemm(M, C, 4 * C, res=True, count=cnt, name="ff2")
    tot["attn"] += bench_attn(B2, 8, HW, HW, C // 8, count=cnt, name="self")
    tot["attn"] += bench_attn(B2, 8, HW, 77, C // 8, count=cnt, name="cross")
    tot["ln"] += bench_ln(M, C, count=3 * cnt)
    tot["gn"] += bench_gn(B2, HW, C, count=cnt, silu=False)
# 1x1 shortcuts
for HW, cin, cout, cnt in ((1024, 320, 640, 1), (256, 640, 1280, 1), (64, 2560, 1280, 3), (256, 2560, 1280, 2), (256, 1920, 1280, 1),
                           (1024, 1920, 640, 1), (1024, 1280, 640, 1), (1024, 960, 640, 1), (4096, 960, 320, 1), (4096, 640, 320, 2)):
    tot["gemm"] += bench_gemm(B2 * HW, cout, cin, count=cnt, name="shortcut")
# convs (SURVEY Appendix B census)
for H, cin, cout, s, cnt in ((64, 320, 320, 1, 7), (64, 640, 320, 1, 2), (64, 960, 320, 1, 1), (32, 640, 640, 1, 6), (32, 320, 640, 1, 1),
                             (32, 960, 640, 1, 1), (32, 1280, 640, 1, 1), (32, 1920, 640, 1, 1), (16, 1280, 1280, 1, 6),
                             (16, 640, 1280, 1, 1), (16, 1920, 1280, 1, 1), (16, 2560, 1280, 1, 2), (8, 1280, 1280, 1, 11),
                             (8, 2560, 1280, 1, 3), (64, 320, 320, 2, 1), (32, 640, 640, 2, 1), (16, 1280, 1280, 2, 1),
                             (16, 1280, 1280, 1, 1), (32, 1280, 1280, 1, 1), (64, 640, 640, 1, 1)):
    tot["conv"] += bench_conv(B2, H, H, cin, cout, s, count=cnt, name="conv")
# group norms of the resnets
for HW, C, cnt in ((4096, 320, 9), (4096, 640, 2), (4096, 960, 1), (1024, 640, 8), (1024, 320, 1), (1024, 960, 1), (1024, 1280, 1),
                   (1024, 1920, 1), (256, 1280, 8), (256, 640, 1), (256, 1920, 1), (256, 2560, 2), (64, 1280, 14), (64, 2560, 3)):
    tot["gn"] += bench_gn(B2, HW, C, count=cnt)
print("per-forward totals (us):", {k: round(v) for k, v in tot.items()}, "sum", round(sum(tot.values())))
ideal_tot = {}
for kind, name, us, tf, gb, ideal, frac in rows:
    cnt = int(name.rsplit("x", 1)[1])
    ideal_tot[kind] = ideal_tot.get(kind, 0.0) + ideal * cnt
print("per-forward ideal (us):", {k: round(v) for k, v in ideal_tot.items()}, "sum", round(sum(ideal_tot.values())))
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"rows": rows, "totals_us": tot, "ideal_us": ideal_tot}, open("gpurun_out/kernel_shapes.json", "w"), indent=1)
