"""Per-kernel parity: every building-block CUDA kernel, called through the C ABI, against a plain PyTorch
fp32 evaluation of the same op on the same (fp16-rounded) inputs.  Tolerances are stated per test: the
kernels accumulate in fp32 and round once to fp16 on output, so the bound is fp16 output rounding
(2^-11 relative) plus accumulation-order noise."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    from gyre_b200 import _native
    _native.load()
    return _native


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float16):
    g = torch.Generator("cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


def assert_close(got, ref, atol, rtol, what):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} / {bad.numel()} out of tolerance, max abs err {err.max().item():.4g} " \
                          f"(ref max {ref.abs().max().item():.4g}) first bad idx {bad.nonzero()[0].tolist()}"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 320, 320), (100, 64, 128), (4096, 320, 320), (300, 1280, 1280),
                                   (77 * 2, 640, 768), (16, 1280, 320), (512, 4, 320), (130, 160, 72)])
def test_gemm_shapes(nat, M, N, K):
    a = rnd(M, K, seed=1)
    w = rnd(N, K, seed=2, scale=1 / math.sqrt(K))
    bias = rnd(N, seed=3, dtype=torch.float32)
    out = nat.gemm(a, w, bias=bias)
    ref = a.float() @ w.float().t() + bias
    assert_close(out, ref, 2e-3, 2e-3, f"gemm {M}x{N}x{K}")


@pytest.mark.parametrize("mode", [3, 1])
@pytest.mark.parametrize("M,N,K", [(4096, 320, 320), (300, 1280, 1280), (129, 640, 2560), (1000, 2560, 320), (128, 256, 64),
                                   (65536, 320, 320), (16384, 1920, 640)])
def test_gemm_pair_mode_bit_identical(nat, M, N, K, mode):
    """CTA-pair modes (2: one tcgen05.mma.cta_group::2 of M = 256 per k step, each CTA holding half of the B tile;
    1: B tiles multicast to both CTAs) change who loads and issues what, not the arithmetic: outputs are bit-identical
    to the single-CTA kernel, including an odd number of M tiles (phantom tile)."""
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=1 / math.sqrt(K))
    bias = rnd(N, seed=3, dtype=torch.float32)
    res = rnd(M, N, seed=4)
    old, old_sk = nat.get_tunable("MCAST"), nat.get_tunable("STREAMK")
    try:
        nat.set_tunable("STREAMK", 0)           # stream-K changes the fp32 summation order
        nat.set_tunable("MCAST", mode)
        o1 = nat.gemm(a, w, bias=bias, residual=res)
        nat.set_tunable("MCAST", 0)
        o0 = nat.gemm(a, w, bias=bias, residual=res)
    finally:
        nat.set_tunable("MCAST", old)
        nat.set_tunable("STREAMK", old_sk)
    assert torch.equal(o0, o1)
    assert_close(o1, a.float() @ w.float().t() + bias + res.float(), 2e-3, 2e-3, "gemm pair mode")


@pytest.mark.parametrize("M,N,K,res", [(4096, 1280, 10240, True), (16384, 640, 5760, True), (1024, 1280, 11520, False),
                                         (4096, 3840, 6400, False), (2000, 1280, 7680, True), (4096, 1280, 1280, True),
                                         (2100, 1280, 7680, True)])
def test_gemm_stream_k(nat, M, N, K, res):
    """Stream-K over the partial last wave: same result as whole-tile scheduling up to fp32 summation order
    (partials are added in a fixed slot order, so repeated runs are bit-identical)."""
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=1 / math.sqrt(K))
    bias = rnd(N, seed=3, dtype=torch.float32)
    r = rnd(M, N, seed=4) if res else None
    old = nat.get_tunable("STREAMK")
    try:
        nat.set_tunable("STREAMK", 1)
        o1 = nat.gemm(a, w, bias=bias, residual=r)
        o1b = nat.gemm(a, w, bias=bias, residual=r)
        nat.set_tunable("STREAMK", 0)
        o0 = nat.gemm(a, w, bias=bias, residual=r)
    finally:
        nat.set_tunable("STREAMK", old)
    assert torch.equal(o1, o1b), "stream-K must be deterministic"
    ref = a.float() @ w.float().t() + bias + (r.float() if res else 0)
    assert_close(o1, ref, 2e-3, 2e-3, "gemm stream-K")
    assert (o1.float() - o0.float()).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


def test_conv_stream_k(nat):
    for (B, H, W, Cin, Cout) in [(16, 8, 8, 1280, 1280), (16, 16, 16, 1280, 1280), (16, 32, 32, 640, 640), (4, 8, 8, 2560, 1280)]:
        x = rnd(B, H, W, Cin, seed=1)
        w = rnd(Cout, Cin, 3, 3, seed=2, scale=1 / math.sqrt(9 * Cin))
        bias = rnd(Cout, seed=3, dtype=torch.float32)
        temb = rnd(B, Cout, seed=4)
        res = rnd(B * H * W, Cout, seed=5)
        wp = nat.pack_conv3x3(w)
        old = nat.get_tunable("STREAMK")
        try:
            nat.set_tunable("STREAMK", 1)
            o1 = nat.conv3x3(x, wp, Cout, bias=bias, residual=res, rowgroup_bias=temb, rows_per_group=H * W)
            nat.set_tunable("STREAMK", 0)
            o0 = nat.conv3x3(x, wp, Cout, bias=bias, residual=res, rowgroup_bias=temb, rows_per_group=H * W)
        finally:
            nat.set_tunable("STREAMK", old)
        ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, padding=1) + temb.float()[:, :, None, None]
        ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, Cout) + res.float()
        assert_close(o1.reshape(B * H * W, Cout), ref, 4e-3, 4e-3, f"conv stream-K {B}x{H}x{W} {Cin}->{Cout}")
        assert (o1.float() - o0.float()).abs().max().item() < 4e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("mode", [3, 1])
def test_conv_pair_mode_bit_identical(nat, mode):
    for (B, H, W, Cin, Cout, stride) in [(2, 32, 32, 320, 320, 1), (3, 8, 8, 1280, 1280, 1), (2, 32, 32, 320, 320, 2),
                                         (1, 24, 24, 192, 160, 1), (16, 32, 32, 640, 640, 1)]:
        x = rnd(B, H, W, Cin, seed=1)
        w = rnd(Cout, Cin, 3, 3, seed=2, scale=1 / math.sqrt(9 * Cin))
        bias = rnd(Cout, seed=3, dtype=torch.float32)
        wp = nat.pack_conv3x3(w)
        old, old_sk = nat.get_tunable("MCAST"), nat.get_tunable("STREAMK")
        try:
            nat.set_tunable("STREAMK", 0)
            nat.set_tunable("MCAST", mode)
            o1 = nat.conv3x3(x, wp, Cout, bias=bias, stride=stride)
            nat.set_tunable("MCAST", 0)
            o0 = nat.conv3x3(x, wp, Cout, bias=bias, stride=stride)
        finally:
            nat.set_tunable("MCAST", old)
            nat.set_tunable("STREAMK", old_sk)
        assert torch.equal(o0, o1), f"conv {B}x{H}x{W} {Cin}->{Cout} s{stride}"


def test_gemm_residual_f32_out(nat):
    M, N, K = 384, 320, 1280
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=1 / math.sqrt(K))
    res = rnd(M, N, seed=4)
    out = nat.gemm(a, w, residual=res)
    ref = a.float() @ w.float().t() + res.float()
    assert_close(out, ref, 2e-3, 2e-3, "gemm+residual")
    out32 = nat.gemm(a, w, out_dtype=torch.float32)
    assert_close(out32, a.float() @ w.float().t(), 2e-4, 1e-4, "gemm f32 out")


def test_gemm_two_sources(nat):
    M, K1, K2, N = 300, 128, 64, 192
    a1, a2 = rnd(M, K1, seed=1), rnd(M, K2, seed=5)
    w = rnd(N, K1 + K2, seed=2, scale=0.1)
    out = nat.gemm(a1, w, a2=a2)
    ref = torch.cat([a1, a2], 1).float() @ w.float().t()
    assert_close(out, ref, 2e-3, 2e-3, "gemm two-source")


@pytest.mark.parametrize("fast", [1, 0])
def test_gemm_geglu(nat, fast):
    """GEGLU epilogue with the branch-free erf (tunable GELU_FAST=1, |erf err| <= 1.5e-7) and with libdevice erff."""
    M, C = 200, 128
    a = rnd(M, C, seed=1, scale=2.0)
    w = rnd(8 * C, C, seed=2, scale=1 / math.sqrt(C))
    b = rnd(8 * C, seed=3, dtype=torch.float32)
    wp, bp = nat.pack_geglu(w, b)
    old = nat.get_tunable("GELU_FAST")
    try:
        nat.set_tunable("GELU_FAST", fast)
        out = nat.gemm(a, wp, bias=bp, act=1)
    finally:
        nat.set_tunable("GELU_FAST", old)
    h = a.float() @ w.float().t() + b
    val, gate = h.chunk(2, dim=-1)
    ref = val * F.gelu(gate)
    assert_close(out, ref, 3e-3, 3e-3, "geglu")


def test_gemm_rowgroup_bias_silu(nat):
    M, N, K = 256, 64, 64
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.125)
    rgb = rnd(4, N, seed=7)
    out = nat.gemm(a, w, rowgroup_bias=rgb, rows_per_group=64)
    ref = a.float() @ w.float().t() + rgb.float().repeat_interleave(64, 0)
    assert_close(out, ref, 2e-3, 2e-3, "rowgroup bias")
    out = nat.gemm(a, w, act=2)
    assert_close(out, F.silu(a.float() @ w.float().t()), 2e-3, 2e-3, "silu epilogue")


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,pad", [
    (1, 16, 16, 64, 64, 1, 1), (2, 8, 8, 128, 96, 1, 1), (2, 64, 64, 320, 320, 1, 1), (1, 32, 32, 640, 320, 1, 1),
    (3, 8, 8, 1280, 1280, 1, 1), (2, 16, 16, 64, 64, 2, 1), (2, 32, 32, 320, 320, 2, 1), (1, 16, 16, 128, 128, 2, 0),
    (1, 64, 64, 320, 4, 1, 1), (1, 12, 20, 64, 32, 1, 1), (2, 24, 24, 192, 160, 1, 1), (1, 96, 96, 128, 128, 1, 1),
])
def test_conv3x3(nat, B, H, W, Cin, Cout, stride, pad):
    x = rnd(B, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, 3, 3, seed=2, scale=1 / math.sqrt(9 * Cin))
    bias = rnd(Cout, seed=3, dtype=torch.float32)
    wp = nat.pack_conv3x3(w)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    out = nat.conv3x3(x_nhwc, wp, Cout, bias=bias, stride=stride, pad=pad)
    xin = x.float()
    if pad == 0:
        xin = F.pad(xin, (0, 1, 0, 1))
    ref = F.conv2d(xin, w.float(), bias, stride=stride, padding=pad).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    assert_close(out, ref, 3e-3, 3e-3, f"conv {B}x{H}x{W} {Cin}->{Cout} s{stride} p{pad}")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 8, 64, 64), (2, 16, 16, 128, 96), (2, 32, 32, 640, 640),
                                           (1, 12, 20, 64, 32), (3, 8, 8, 1280, 1280), (1, 64, 64, 256, 256)])
def test_upsample_conv_folded(nat, B, H, W, Cin, Cout):
    """conv3x3(nearest-2x(x)) as four 2x2 phase convs on the low-res input vs the un-folded fp32 reference.
    The phase weights are sums of up to four fp16 taps rounded once to fp16: slightly wider tolerance."""
    x = rnd(B, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, 3, 3, seed=2, scale=1 / math.sqrt(9 * Cin))
    bias = rnd(Cout, seed=3, dtype=torch.float32)
    wp4 = nat.pack_upconv3x3(w)
    out = nat.upconv2x(x.permute(0, 2, 3, 1).contiguous(), wp4, Cout, bias=bias)
    up = F.interpolate(x.float(), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w.float(), bias, padding=1).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    assert_close(out, ref, 4e-3, 4e-3, f"upconv {B}x{H}x{W} {Cin}->{Cout}")


def test_upsample_conv_matches_unfolded_kernel(nat):
    """Same op through the un-folded kernels (upsample handled by the caller): the two paths agree to fp16 noise."""
    B, H, W, C = 2, 16, 16, 320
    x = rnd(B, C, H, W, seed=5)
    w = rnd(C, C, 3, 3, seed=6, scale=1 / math.sqrt(9 * C))
    bias = rnd(C, seed=7, dtype=torch.float32)
    folded = nat.upconv2x(x.permute(0, 2, 3, 1).contiguous(), nat.pack_upconv3x3(w), C, bias=bias)
    up = F.interpolate(x, scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1).contiguous()
    plain = nat.conv3x3(up, nat.pack_conv3x3(w), C, bias=bias)
    assert_close(folded, plain, 4e-3, 4e-3, "folded vs plain upsample conv")


def test_conv3x3_fused_epilogue(nat):
    B, H, W, Cin, Cout = 2, 16, 16, 64, 128
    x = rnd(B, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, 3, 3, seed=2, scale=1 / math.sqrt(9 * Cin))
    bias = rnd(Cout, seed=3, dtype=torch.float32)
    temb = rnd(B, Cout, seed=4)
    res = rnd(B * H * W, Cout, seed=5)
    wp = nat.pack_conv3x3(w)
    out = nat.conv3x3(x.permute(0, 2, 3, 1).contiguous(), wp, Cout, bias=bias, residual=res, rowgroup_bias=temb,
                      rows_per_group=H * W)
    ref = F.conv2d(x.float(), w.float(), bias, padding=1) + temb.float()[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, Cout) + res.float()
    assert_close(out.reshape(B * H * W, Cout), ref, 3e-3, 3e-3, "conv fused epilogue")


@pytest.mark.parametrize("B,HW,C1,C2,G,silu", [(2, 256, 64, 0, 32, True), (2, 4096, 320, 0, 32, True),
                                               (1, 1024, 640, 320, 32, True), (3, 64, 1280, 1280, 32, False),
                                               (1, 100, 128, 0, 32, False), (1, 65536, 128, 0, 32, True),
                                               (2, 256, 1280, 640, 32, True), (2, 1000, 960, 0, 32, True),
                                               (1, 4096, 640, 320, 32, True), (2, 64, 2560, 0, 32, True),
                                               (1, 300, 96, 32, 16, True)])
def test_groupnorm(nat, B, HW, C1, C2, G, silu):
    x1 = rnd(B, HW, C1, seed=1) + 0.5
    x2 = rnd(B, HW, C2, seed=2, scale=2.0) if C2 else None
    C = C1 + C2
    gamma = 1 + 0.1 * rnd(C, seed=3, dtype=torch.float32)
    beta = 0.1 * rnd(C, seed=4, dtype=torch.float32)
    out = nat.groupnorm(x1, gamma, beta, G, 1e-5, silu, x2=x2)
    xin = x1 if x2 is None else torch.cat([x1, x2], dim=2)
    ref = F.group_norm(xin.float().permute(0, 2, 1), G, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    assert_close(out, ref.permute(0, 2, 1), 2e-3, 2e-3, "groupnorm")


# shapes: UNet levels 0 / 1 (cpg 10 / 20: groups straddle the 32-column epilogue slices), SDXL-like 1280 at 32x32 (cpg 40),
# the stride-2 downsamplers, VAE widths (cpg 4 / 8 / 16; many partials -> the fold launch), stream-K and pair launches
@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,pad,res,temb", [
    (2, 64, 64, 320, 320, 1, 1, True, False), (8, 32, 32, 640, 640, 1, 1, False, True), (8, 32, 32, 320, 640, 1, 1, False, True),
    (16, 32, 32, 128, 1280, 1, 1, True, False), (2, 128, 128, 320, 320, 2, 1, False, False), (16, 64, 64, 64, 640, 2, 1, False, False),
    (1, 128, 128, 256, 256, 1, 1, False, False), (1, 64, 64, 512, 512, 1, 1, True, False),
    (1, 256, 256, 128, 128, 1, 1, False, False), (16, 64, 64, 64, 320, 1, 1, False, True), (4, 128, 128, 64, 128, 2, 0, False, False)])
def test_conv3x3_groupnorm_statistics(nat, B, H, W, Cin, Cout, stride, pad, res, temb):
    """The conv epilogue's GroupNorm partials (Epilogue::gn_out) + groupnorm_pre against (a) fp64 statistics of the conv's own
    fp16 output and (b) the two-pass GroupNorm kernels on that output."""
    G = 32
    x = rnd(B, H, W, Cin, seed=1)
    w = rnd(Cout, Cin, 3, 3, seed=2, scale=1 / math.sqrt(9 * Cin))
    bias = rnd(Cout, seed=3, dtype=torch.float32)
    wp = nat.pack_conv3x3(w)
    parts = nat.load().gyre_b200_conv3x3_gn_parts(B, H, W, Cout, stride, pad, G)
    assert parts > 0, "this shape was chosen to take the fused path"
    Ho, Wo = (H, W) if stride == 1 else (((H - 1) // 2 + 1, (W - 1) // 2 + 1) if pad == 1 else ((H - 2) // 2 + 1, (W - 2) // 2 + 1))
    residual = (rnd(B * Ho * Wo, Cout, seed=4) + 0.25) if res else None
    tvec = rnd(B, Cout, seed=5) if temb else None
    kw = dict(bias=bias, residual=residual, stride=stride, pad=pad, rowgroup_bias=tvec, rows_per_group=Ho * Wo)
    out, pre = nat.conv3x3(x, wp, Cout, gn_groups=G, **kw)
    # producing the statistics does not change the convolution's output: bitwise with stream-K off (with it on, the plain
    # launch of a small map may take a wider, K-split tile, i.e. another fp32 summation order)
    old_sk = nat.get_tunable("STREAMK")
    nat.set_tunable("STREAMK", 0)
    try:
        assert torch.equal(nat.conv3x3(x, wp, Cout, gn_groups=G, **kw)[0], nat.conv3x3(x, wp, Cout, **kw))
    finally:
        nat.set_tunable("STREAMK", old_sk)
    assert_close(out, nat.conv3x3(x, wp, Cout, **kw), 2e-3, 2e-3, "conv with / without statistics")
    assert pre.shape == (B, parts, G, 2) and torch.isfinite(pre).all(), "every (sample, tile, group) partial is written"
    o64 = out.double().view(B, Ho * Wo, G, Cout // G)
    tot = pre.double().sum(1)
    n = Ho * Wo * (Cout // G)
    assert_close(tot[..., 0] / n, o64.sum((1, 3)) / n, 1e-5, 1e-5, "group means")
    assert_close(tot[..., 1] / n, (o64 * o64).sum((1, 3)) / n, 1e-5, 1e-5, "group second moments")
    gamma = 1 + 0.1 * rnd(Cout, seed=6, dtype=torch.float32)
    beta = 0.1 * rnd(Cout, seed=7, dtype=torch.float32)
    assert nat.load().gyre_b200_groupnorm_pre_ok(Cout, Ho * Wo, G) == 1
    flat = out.view(B, Ho * Wo, Cout)
    for silu in (True, False):
        fused = nat.groupnorm_pre(flat, gamma, beta, G, 1e-5, silu, pre)
        two_pass = nat.groupnorm(flat, gamma, beta, G, 1e-5, silu)
        ref = F.group_norm(flat.float().permute(0, 2, 1), G, gamma, beta, 1e-5)
        ref = (F.silu(ref) if silu else ref).permute(0, 2, 1)
        assert_close(fused, ref, 2e-3, 2e-3, "groupnorm from conv statistics")
        assert (fused.float() - two_pass.float()).abs().max().item() <= 2e-3
    # batch independence: a sample's partials (and so its normalised rows) are bit-identical at batch 1
    one, pre1 = nat.conv3x3(x[B - 1:].contiguous(), wp, Cout, gn_groups=G, bias=bias, stride=stride, pad=pad,
                            residual=residual[(B - 1) * Ho * Wo:].contiguous() if res else None,
                            rowgroup_bias=tvec[B - 1:].contiguous() if temb else None, rows_per_group=Ho * Wo)
    if pre1 is not None and torch.equal(one[0], out[B - 1]):      # (stream-K may change the conv's last bit with the batch)
        assert torch.equal(pre1[0], pre[B - 1])


def test_conv3x3_groupnorm_statistics_odd_tile_count(nat):
    """9 M tiles (3 per sample, batch 3) under CTA pairs: the last pair's second tile does not exist - nothing may be written
    for it (guard sample stays NaN), every real partial is written."""
    B, H, W, Cin, Cout, G = 3, 16, 24, 128, 128, 32
    x = rnd(B, H, W, Cin, seed=1)
    w = rnd(Cout, Cin, 3, 3, seed=2, scale=1 / math.sqrt(9 * Cin))
    wp = nat.pack_conv3x3(w)
    old = nat.get_tunable("FORCE_BN")
    nat.set_tunable("FORCE_BN", 128)        # (a problem this small would pick 64-wide tiles, which carry no statistics)
    try:
        assert nat.load().gyre_b200_conv3x3_gn_parts(B, H, W, Cout, 1, 1, G) == 3
        out, pre = nat.conv3x3(x, wp, Cout, gn_groups=G, gn_guard=True)
    finally:
        nat.set_tunable("FORCE_BN", old)
    assert pre.shape == (B + 1, 3, G, 2)
    assert torch.isnan(pre[B]).all(), "the phantom tile of the last CTA pair wrote statistics"
    assert torch.isfinite(pre[:B]).all()
    o64 = out.double().view(B, H * W, G, Cout // G)
    assert_close(pre[:B].double().sum(1)[..., 0], o64.sum((1, 3)), 1e-2, 1e-5, "group sums")


def test_conv3x3_gn_parts_refuses_what_it_cannot_do(nat):
    lib = nat.load()
    assert lib.gyre_b200_conv3x3_gn_parts(2, 8, 8, 1280, 1, 1, 32) == 0       # a tile spans two samples
    assert lib.gyre_b200_conv3x3_gn_parts(1, 12, 20, 320, 1, 1, 32) == 0      # tiles overhang the image
    assert lib.gyre_b200_conv3x3_gn_parts(1, 64, 64, 96, 1, 1, 32) == 0       # odd-width groups (cpg 3)
    assert lib.gyre_b200_conv3x3_gn_parts(1, 64, 64, 4, 1, 1, 32) == 0
    assert lib.gyre_b200_conv3x3_gn_parts(2, 64, 64, 320, 1, 1, 32) == 32
    x = rnd(1, 12, 20, 64, seed=1)
    w = rnd(320, 64, 3, 3, seed=2, scale=0.05)
    out, pre = nat.conv3x3(x, nat.pack_conv3x3(w), 320, gn_groups=32)
    assert pre is None and out.shape == (1, 12, 20, 320)


def test_groupnorm_batch_independent(nat):
    """Chunking depends on HW only: a sample's result is bit-identical at any batch size."""
    x = rnd(4, 4096, 320, seed=9)
    gamma = 1 + 0.1 * rnd(320, seed=3, dtype=torch.float32)
    beta = 0.1 * rnd(320, seed=4, dtype=torch.float32)
    full = nat.groupnorm(x, gamma, beta, 32, 1e-5, True)
    one = nat.groupnorm(x[2:3].contiguous(), gamma, beta, 32, 1e-5, True)
    assert torch.equal(full[2:3], one)


@pytest.mark.parametrize("sub", [2, 1, 0])
@pytest.mark.parametrize("rows,C", [(77, 64), (4096, 320), (1000, 640), (257, 1280), (1, 320), (4099, 320), (33, 8), (61, 768),
                                    (130, 576)])
def test_layernorm(nat, rows, C, sub):
    """sub = 1 / 2: several rows per warp for C <= 320 / 640 (ragged last warp / CTA included); 0: one warp per row."""
    x = rnd(rows, C, seed=1) * 2 + 0.3
    gamma = 1 + 0.1 * rnd(C, seed=3, dtype=torch.float32)
    beta = 0.1 * rnd(C, seed=4, dtype=torch.float32)
    old = nat.get_tunable("LN_SUB")
    try:
        nat.set_tunable("LN_SUB", sub)
        out = nat.layernorm(x, gamma, beta)
        # a row's result does not depend on the rows around it (batch invariance)
        one = nat.layernorm(x[rows // 2:rows // 2 + 1].contiguous(), gamma, beta)
    finally:
        nat.set_tunable("LN_SUB", old)
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    assert_close(out, ref, 2e-3, 2e-3, "layernorm")
    assert torch.equal(one[0], out[rows // 2])


@pytest.mark.parametrize("n", [1, 1000, 2 * 4 * 64 * 64, 300001])
def test_dpm_error_norm(nat, n):
    """Error estimate of the adaptive DPM-Solver (k_diffusion/sampling.py:461-462) against its definition."""
    import ctypes as C
    g = torch.Generator().manual_seed(n)
    lo = torch.randn(n, generator=g).cuda()
    hi = (lo.cpu() + 0.01 * torch.randn(n, generator=g)).cuda()
    prev = torch.randn(n, generator=g).cuda() * 3
    atol, rtol = 0.0078, 0.05
    lib = nat.load()
    parts = torch.empty(lib.gyre_b200_dpm_error_num_partials(), device="cuda", dtype=torch.float64)
    nat.check(lib.gyre_b200_dpm_error_partials(nat.ptr(lo), nat.ptr(hi), nat.ptr(prev), atol, rtol, n, nat.ptr(parts),
                                               nat.stream_ptr()), "dpm_error_partials")
    got = math.sqrt(sum(parts.cpu().tolist())) / n ** 0.5
    delta = torch.maximum(torch.tensor(atol), rtol * torch.maximum(lo.cpu().abs(), prev.cpu().abs()))
    ref = float(torch.linalg.norm(((lo.cpu() - hi.cpu()) / delta).double()) / n ** 0.5)
    assert abs(got - ref) <= 1e-5 * max(ref, 1e-6), (got, ref)
    parts2 = torch.empty_like(parts)
    nat.check(lib.gyre_b200_dpm_error_partials(nat.ptr(lo), nat.ptr(hi), nat.ptr(prev), atol, rtol, n, nat.ptr(parts2),
                                               nat.stream_ptr()), "dpm_error_partials")
    assert torch.equal(parts, parts2), "the partial sums must be reproducible"


@pytest.mark.parametrize("M,C,N2", [(4096, 320, 960), (65536, 320, 320), (16384, 640, 1920), (4096, 1280, 1280), (1000, 320, 320),
                                    (300, 64, 192)])
def test_layernorm_folded_into_gemms(nat, M, C, N2):
    """The producing GEMM leaves per-row (sum, sum of squares) partials of what it writes; the consuming GEMM reads the RAW
    rows against gamma-scaled weights and applies rstd * (acc - mean * colsum) + bias' in its epilogue.  Reference:
    F.layer_norm of the produced fp16 tensor followed by F.linear, in fp32."""
    a = rnd(M, C, seed=1)
    w1 = rnd(C, C, seed=2, scale=1 / math.sqrt(C))
    b1 = rnd(C, seed=3, dtype=torch.float32)
    res = rnd(M, C, seed=4, scale=2.0) + 0.5          # a residual stream with a non-zero row mean
    gamma = (1.0 + 0.3 * rnd(C, seed=5, dtype=torch.float32)).contiguous()
    beta = 0.2 * rnd(C, seed=6, dtype=torch.float32)
    w2 = rnd(N2, C, seed=7, scale=1 / math.sqrt(C))
    b2 = rnd(N2, seed=8, dtype=torch.float32)
    parts = nat.gemm_rowstat_buffer(M, C, a.device)
    h = nat.gemm(a, w1, bias=b1, residual=res, rowstat_out=parts)
    href = a.float() @ w1.float().t() + b1 + res.float()
    assert_close(h, href, 4e-3, 2e-3, "producer output")
    # the statistics are those of the rounded tensor up to the rounding itself
    stat = nat.ln_finalize_rows(parts, C)
    hf = h.float()
    mean_ref = hf.mean(dim=1)
    rstd_ref = 1.0 / torch.sqrt(hf.var(dim=1, unbiased=False) + 1e-5)
    assert (stat[:, 0] - mean_ref).abs().max().item() < 2e-3
    assert ((stat[:, 1] - rstd_ref).abs() / rstd_ref).max().item() < 2e-3
    wf, cs, lb = nat.ln_fold_linear(w2, gamma, beta, b2)
    out = nat.gemm(h, wf, bias=lb, ln_rowstat=stat, ln_colsum=cs)
    ref = F.linear(F.layer_norm(hf, (C,), gamma, beta, 1e-5), w2.float(), b2)
    assert_close(out, ref, 6e-3, 4e-3, f"folded LayerNorm -> Linear {M}x{C}->{N2}")
    # the unfused pair of kernels is no closer to the fp32 reference than the folded form
    ln = nat.layernorm(h, gamma, beta)
    unfused = nat.gemm(ln, w2, bias=b2)
    if parts.shape[0] <= 8:
        # narrow rows: the consumer folds the raw partials itself, no finalize launch
        out_direct = nat.gemm(h, wf, bias=lb, ln_rowstat=parts, ln_colsum=cs, ln_raw_parts=True)
        assert_close(out_direct, ref, 6e-3, 4e-3, f"folded LayerNorm (raw partials) -> Linear {M}x{C}->{N2}")
        assert (out_direct.float() - out.float()).abs().max().item() < 4e-3
    e_f = (out.float() - ref).abs().max().item()
    e_u = (unfused.float() - ref).abs().max().item()
    print(f"folded LN {M}x{C}->{N2}: max abs err folded {e_f:.3e}, LayerNorm kernel + GEMM {e_u:.3e}")
    assert e_f <= 1.5 * e_u + 1e-3


@pytest.mark.parametrize("M,C", [(4096, 320), (16384, 640), (1024, 1280)])
def test_layernorm_folded_into_geglu(nat, M, C):
    a = rnd(M, C, seed=1)
    w1 = rnd(C, C, seed=2, scale=1 / math.sqrt(C))
    res = rnd(M, C, seed=4) - 0.3
    gamma = (1.0 + 0.3 * rnd(C, seed=5, dtype=torch.float32)).contiguous()
    beta = 0.2 * rnd(C, seed=6, dtype=torch.float32)
    Fdim = 4 * C
    w2 = rnd(2 * Fdim, C, seed=7, scale=1 / math.sqrt(C))
    b2 = rnd(2 * Fdim, seed=8, dtype=torch.float32)
    parts = nat.gemm_rowstat_buffer(M, C, a.device)
    h = nat.gemm(a, w1, residual=res, rowstat_out=parts)
    stat = nat.ln_finalize_rows(parts, C)
    wp, bp = nat.pack_geglu(w2, b2)
    wf, cs, lb = nat.ln_fold_linear(wp, gamma, beta, bp)
    out = nat.gemm(h, wf, bias=lb, act=1, ln_rowstat=stat, ln_colsum=cs)
    y = F.linear(F.layer_norm(h.float(), (C,), gamma, beta, 1e-5), w2.float(), b2)
    ref = y[:, :Fdim] * F.gelu(y[:, Fdim:])
    assert_close(out, ref, 8e-3, 6e-3, f"folded LayerNorm -> GEGLU {M}x{C}")
