"""Attention kernel (tcgen05 flash attention) through the C ABI against fp32 softmax(QK^T)V in PyTorch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    from gyre_b200 import _native
    _native.load()
    return _native


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator("cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).half().cuda()


def ref_attention(q, k, v, heads):
    B, Nq, C = q.shape
    Nk = k.shape[1]
    d = C // heads
    qh = q.float().view(B, Nq, heads, d).permute(0, 2, 1, 3)
    kh = k.float().view(B, Nk, heads, d).permute(0, 2, 1, 3)
    vh = v.float().view(B, Nk, heads, d).permute(0, 2, 1, 3)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, dim=-1)
    return (p @ vh).permute(0, 2, 1, 3).reshape(B, Nq, C)


# (B, heads, Nq, Nk, d): SD1.5 levels (d=40/80/160), SD2 (d=64), cross (Nk=77), ToMe-merged (Nk=N-r), tiny dims
CASES = [(1, 1, 128, 128, 64), (2, 4, 256, 256, 16), (2, 8, 1024, 1024, 40), (1, 8, 256, 256, 80),
         (2, 8, 64, 64, 160), (2, 5, 300, 300, 64), (2, 8, 1024, 77, 40), (1, 8, 256, 77, 160), (1, 2, 100, 13, 32),
         (1, 8, 1024, 768, 40), (1, 8, 4096, 4096, 40), (1, 4, 256, 130, 80),
         # short-key kernel (Nk <= 128, Nq >= 512): cross-attention at SD1.5 / SD2.1 / SDXL shapes, ragged edges
         (2, 8, 4096, 77, 40), (2, 8, 1024, 77, 80), (1, 10, 2304, 77, 64), (1, 5, 9216, 77, 64),
         (1, 3, 700, 77, 40), (2, 4, 512, 128, 64), (1, 2, 640, 16, 128), (1, 8, 1000, 100, 72),
         # ping-pong kernel at 64 < d <= 128 (P aliased onto S in TMEM): SD1.5 level 1, ragged keys, d = 128
         (2, 8, 1024, 1024, 80), (1, 5, 640, 640, 128), (1, 4, 384, 300, 96), (1, 2, 2048, 1500, 104), (1, 8, 300, 290, 80),
         # cross-attention against LPW multi-chunk prompts (2 / 3 chunks of 77 tokens) at the SD1.5 / SD2.1 levels
         (2, 8, 4096, 154, 40), (2, 8, 4096, 231, 40), (1, 8, 1024, 231, 80), (2, 8, 256, 231, 160), (1, 8, 64, 231, 160),
         (1, 10, 2304, 231, 64)]


@pytest.mark.parametrize("B,heads,Nq,Nk,d", CASES)
def test_attention(nat, B, heads, Nq, Nk, d):
    C = heads * d
    q, k, v = rnd(B, Nq, C, seed=1), rnd(B, Nk, C, seed=2), rnd(B, Nk, C, seed=3)
    out = nat.attention(q, k, v, heads)
    ref = ref_attention(q, k, v, heads)
    err = (out.float() - ref).abs().max().item()
    # P is rounded to fp16 before P.V and the output to fp16: 2^-11 relative on O(1) values
    assert err < 4e-3, f"attention B{B} h{heads} Nq{Nq} Nk{Nk} d{d}: max abs err {err}"


def test_attention_fused_qkv_views(nat):
    """q/k/v as column slices of one fused projection output (how the UNet calls it)."""
    B, N, heads, d = 2, 256, 8, 40
    C = heads * d
    qkv = rnd(B, N, 3 * C, seed=5)
    out = nat.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads)
    ref = ref_attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads)
    assert (out.float() - ref).abs().max().item() < 4e-3


def test_attention_peaked_softmax(nat):
    """Large logits exercise the lazy-rescale path (running max grows by > 2^8 between key tiles)."""
    B, N, heads, d = 1, 512, 2, 64
    C = heads * d
    q, k, v = rnd(B, N, C, seed=1, scale=6.0), rnd(B, N, C, seed=2, scale=6.0), rnd(B, N, C, seed=3)
    out = nat.attention(q, k, v, heads)
    ref = ref_attention(q, k, v, heads)
    assert torch.isfinite(out).all()
    assert (out.float() - ref).abs().max().item() < 2e-2


# softmax variants of the d <= 64 flash kernel (include/gyre_b200.h: tunable "ATT_VARIANT"): 0 plain, 1 staggered
# groups, 3 + packed fp32x2 maths, 7 / 11 + polynomial exp2 on 25 % / 50 % of the scores, 6 packed + poly w/o stagger
@pytest.mark.parametrize("variant", [0, 1, 3, 7, 64, 70, 192, 198, 202, 454, 710, 1006])
def test_attention_softmax_variants(nat, variant):
    old = nat.get_tunable("ATT_VARIANT")
    try:
        nat.set_tunable("ATT_VARIANT", variant)
        for (B, heads, Nq, Nk, d) in [(1, 8, 1024, 1024, 40), (1, 2, 512, 640, 64), (1, 4, 2048, 2000, 40)]:
            C = heads * d
            q, k, v = rnd(B, Nq, C, seed=1, scale=1.5), rnd(B, Nk, C, seed=2, scale=1.5), rnd(B, Nk, C, seed=3)
            out = nat.attention(q, k, v, heads)
            ref = ref_attention(q, k, v, heads)
            err = (out.float() - ref).abs().max().item()
            assert err < 4e-3, f"variant {variant} Nq{Nq} Nk{Nk} d{d}: max abs err {err}"
    finally:
        nat.set_tunable("ATT_VARIANT", old)


def test_attention_optimistic_max_redo_path(nat):
    """Optimistic softmax: logits that keep raising the running max by > 2^8 force the redo path on most tiles."""
    B, N, heads, d = 1, 1024, 2, 64
    C = heads * d
    q, k, v = rnd(B, N, C, seed=1, scale=1.0), rnd(B, N, C, seed=2, scale=1.0), rnd(B, N, C, seed=3)
    # keys grow along the sequence: every later tile dominates the earlier ones
    ramp = torch.linspace(0.2, 12.0, N, device=k.device)[None, :, None]
    k = (k.float() * ramp).half()
    q = (q.float().abs() * 1.5).half()
    old = nat.get_tunable("ATT_VARIANT")
    try:
        for variant in (454, 198, 1006):
            nat.set_tunable("ATT_VARIANT", variant)
            out = nat.attention(q, k, v, heads)
            ref = ref_attention(q, k, v, heads)
            assert torch.isfinite(out).all()
            err = (out.float() - ref).abs().max().item()
            assert err < 2e-2, f"variant {variant}: max abs err {err}"
    finally:
        nat.set_tunable("ATT_VARIANT", old)


def test_attention_d128_kernel_matches_single_tile_kernel(nat):
    B, heads, N, d = 2, 8, 1024, 80
    C = heads * d
    q, k, v = rnd(B, N, C, seed=1), rnd(B, N, C, seed=2), rnd(B, N, C, seed=3)
    old = nat.get_tunable("ATT_D128")
    try:
        nat.set_tunable("ATT_D128", 1)
        a = nat.attention(q, k, v, heads)
        nat.set_tunable("ATT_D128", 0)
        b = nat.attention(q, k, v, heads)
    finally:
        nat.set_tunable("ATT_D128", old)
    assert (a.float() - b.float()).abs().max().item() < 2e-3


def test_attention_short_key_kernel_matches_flash(nat):
    """The K/V-resident kernel and the flash kernel give the same cross-attention to fp16 noise."""
    B, heads, Nq, Nk, d = 2, 8, 4096, 77, 40
    C = heads * d
    q, k, v = rnd(B, Nq, C, seed=1), rnd(B, Nk, C, seed=2), rnd(B, Nk, C, seed=3)
    old = nat.get_tunable("XATTN")
    try:
        nat.set_tunable("XATTN", 1)
        a = nat.attention(q, k, v, heads)
        nat.set_tunable("XATTN", 0)
        b = nat.attention(q, k, v, heads)
    finally:
        nat.set_tunable("XATTN", old)
    assert (a.float() - b.float()).abs().max().item() < 2e-3


# ---- lean-softmax kernel (attention5, ATT_VARIANT 2000 + flags): optimistic exponentials against the running
# reference max with overflow detection on the packed fp16 P words, row sum from a ones column of V
A5_CASES = [(2, 8, 1024, 1024, 40), (1, 8, 4096, 4096, 40), (1, 8, 1024, 768, 40), (2, 5, 300, 300, 64),
            (2, 4, 256, 256, 16), (1, 2, 100, 13, 32), (1, 3, 700, 333, 40), (1, 10, 2304, 2304, 64),
            (1, 4, 512, 640, 56), (1, 8, 256, 77, 40)]


@pytest.mark.parametrize("variant", [2000, 2002, 2004, 2016, 2064, 2130])
@pytest.mark.parametrize("B,heads,Nq,Nk,d", A5_CASES)
def test_attention5(nat, B, heads, Nq, Nk, d, variant):
    C = heads * d
    q, k, v = rnd(B, Nq, C, seed=1), rnd(B, Nk, C, seed=2), rnd(B, Nk, C, seed=3)
    old = nat.get_tunable("ATT_VARIANT")
    nat.set_tunable("ATT_VARIANT", variant)
    try:
        out = nat.attention(q, k, v, heads)
    finally:
        nat.set_tunable("ATT_VARIANT", old)
    err = (out.float() - ref_attention(q, k, v, heads)).abs().max().item()
    assert err < 4e-3, f"attention5 v{variant} B{B} h{heads} Nq{Nq} Nk{Nk} d{d}: max abs err {err}"


@pytest.mark.parametrize("variant", [2000, 2002, 2004, 2016, 2064, 2130])
@pytest.mark.parametrize("kind", ["ramp", "spike", "huge", "late_spike_poly_slot"])
def test_attention5_running_max_growth(nat, variant, kind):
    """Keys whose scores keep growing along the sequence (every tile raises the row max by more than the
    optimistic head-room), single outlier keys, and jumps large enough to overflow fp16 / the polynomial's
    exponent insert: the redo path must give the exact softmax."""
    B, heads, N, d = 1, 2, 1024, 40
    C = heads * d
    g = torch.Generator("cpu").manual_seed(7)
    q = torch.randn(B, N, C, generator=g)
    k = torch.randn(B, N, C, generator=g)
    v = torch.randn(B, N, C, generator=g)
    if kind == "ramp":
        k = k * torch.linspace(0.2, 6.0, N).view(1, N, 1)
    elif kind == "spike":
        k[:, 700] = q[:, 5] * 3.0
        k[:, 130] = q[:, 300] * 2.0
    elif kind == "huge":
        q = q * 6.0
        k = k * 6.0
        k[:, 900] *= 8.0
    else:
        # outlier keys sitting on columns the polynomial path handles (pairs 3, 7, 11, 15 of each 32-column chunk)
        for col in (256 + 6, 256 + 7, 512 + 14, 640 + 30, 896 + 31):
            k[:, col] = q[:, 11] * 40.0
    q, k, v = q.half().cuda(), k.half().cuda(), v.half().cuda()
    old = nat.get_tunable("ATT_VARIANT")
    nat.set_tunable("ATT_VARIANT", variant)
    try:
        out = nat.attention(q, k, v, heads)
    finally:
        nat.set_tunable("ATT_VARIANT", old)
    assert torch.isfinite(out).all()
    err = (out.float() - ref_attention(q, k, v, heads)).abs().max().item()
    assert err < 6e-3, f"attention5 v{variant} {kind}: max abs err {err}"


@pytest.mark.parametrize("name", ["self", "cross", "self_d40", "tome_r16", "tome_half", "tome_odd"])
def test_attention_module_vs_reference_fixture(name):
    """The attention MODULE as the native UNet composes it - projections on the tensor-core GEMM, the ToMe K/V merge, the flash
    kernel, the output projection with its bias - against the output of the reference's own modules
    (MemoryEfficientCrossAttention / ToMeMemoryEfficientCrossAttention; tests/golden/attention.pt)."""
    import os
    from gyre_b200 import _native as N
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "attention.pt"))[name]
    C_, heads, N_, ctx_dim, L, r = g["config"]
    sd = {k: v.cuda() for k, v in g["state_dict"].items()}
    x = g["x"].half().cuda()
    src = x if g["ctx"] is None else g["ctx"].half().cuda()
    B = x.shape[0]
    q = N.gemm(x.reshape(-1, C_), sd["to_q.weight"].half()).reshape(B, N_, C_)
    k = N.gemm(src.reshape(-1, src.shape[-1]), sd["to_k.weight"].half()).reshape(B, src.shape[1], C_)
    v = N.gemm(src.reshape(-1, src.shape[-1]), sd["to_v.weight"].half()).reshape(B, src.shape[1], C_)
    if r:
        k, v = N.tome_merge_kv(k, v, r)
    o = N.attention(q, k, v, heads)
    out = N.gemm(o.reshape(-1, C_), sd["to_out.0.weight"].half(), bias=sd["to_out.0.bias"].float()).reshape(B, N_, C_)
    ref = g["out"]
    err = (out.float().cpu() - ref).abs().max().item()
    # fp16 operands against the fp32 reference module (the ToMe fixtures are built free of near-ties: the same plan in fp16)
    bound = 6e-3 if not r else 1e-2
    assert err < bound * max(1.0, ref.abs().max().item()), f"{name}: max abs err {err} (|ref| <= {ref.abs().max().item():.2f})"
