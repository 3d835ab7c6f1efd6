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
         (1, 8, 1024, 768, 40), (1, 8, 4096, 4096, 40), (1, 4, 256, 130, 80)]


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
