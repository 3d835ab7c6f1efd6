"""gyre_b200.hints wrapper classes (plain torch, run here on the CPU) and the oracle's restatement against what the REFERENCE's
own gyre/pipeline/unet/core.py classes returned over the same fake models (tests/golden/hints.pt, scripts/make_golden.py:pin_hints)."""
import os

import pytest
import torch

from fakes import FakeHintAdapter, FakeHintControlnet, FakeHintUNet

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _impls():
    from gyre_b200 import hints as product
    from oracle import hints as oracle
    return {"product": product, "oracle": oracle}


@pytest.mark.parametrize("impl", ["product", "oracle"])
def test_unet_with_controlnet_matches_reference(impl):
    H = _impls()[impl]
    G = torch.load(os.path.join(GOLD, "hints.pt"))
    I = G["inputs"]
    assert len(G["controlnet"]) == 8
    for key, ref in G["controlnet"].items():
        meta, combo = key.split("/")
        cns = [FakeHintControlnet(100 * (i + 1), c == "c") for i, c in enumerate(combo)]
        got = H.UNetWithControlnet(FakeHintUNet(), cns)(I["lat"], I["t"], encoder_hidden_states=I["ehs"], cfg_meta=meta)
        assert torch.equal(got, ref), key


@pytest.mark.parametrize("impl", ["product", "oracle"])
def test_unet_with_t2i_matches_reference(impl):
    H = _impls()[impl]
    G = torch.load(os.path.join(GOLD, "hints.pt"))
    I = G["inputs"]
    assert len(G["t2i"]) == 12
    for key, ref in G["t2i"].items():
        meta, combo = key.split("/")
        ads = [FakeHintAdapter(1000 * (i + 1), c == "c") for i, c in enumerate(combo)]
        e = torch.cat([I["ehs"][:1], I["ehs"][:1]]) if meta == "f" else I["ehs"][:1]
        l = torch.cat([I["lat"][:1], I["lat"][:1]]) if meta == "f" else I["lat"][:1]
        t = I["t"][:2] if meta == "f" else I["t"][:1]
        w = H.UNetWithT2I(FakeHintUNet(), ads)
        assert torch.equal(w(l, t, encoder_hidden_states=e, cfg_meta=meta), ref), key
        if meta != "u":                                                          # cfg_meta inferred from the batch: "f" or "g"
            assert torch.equal(w(l, t, encoder_hidden_states=e), ref), key


def test_t2i_states_expand_to_the_unet_batch():
    from gyre_b200.hints import UNetWithT2I
    w = UNetWithT2I(None, [FakeHintAdapter(3, False), FakeHintAdapter(5, True)])
    f = w.states_for("f", 6)
    assert [tuple(s.shape) for s in f] == [(6, 4, 8, 8), (6, 8, 4, 4), (6, 12, 2, 2), (6, 16, 1, 1)]
    u, g = w.standard_states["u"], w.standard_states["g"]
    for s, a, b in zip(f, u, g):
        assert torch.equal(s[:3], a.expand(3, -1, -1, -1)) and torch.equal(s[3:], b.expand(3, -1, -1, -1))
    assert w.states_for("f", 6) is f or all(torch.equal(x, y) for x, y in zip(w.states_for("f", 6), f))
    assert [tuple(s.shape) for s in w.states_for("g", 3)] == [(3, 4, 8, 8), (3, 8, 4, 4), (3, 12, 2, 2), (3, 16, 1, 1)]


def test_unbuilt_hint_features_raise():
    from gyre_b200.hints import UNetWithT2I, _split_hint

    class Co(FakeHintAdapter):
        def coadapter_type(self):
            return "sketch"
    with pytest.raises(NotImplementedError):
        UNetWithT2I(None, [Co(1, False)])
    # an RGBA hint carries its mask in the alpha channel; a mask of ones is dropped (unified_pipeline.py:758-771)
    rgba = torch.rand(1, 4, 8, 8)
    img, mask = _split_hint(rgba, None)
    assert torch.equal(img, rgba[:, :3]) and torch.equal(mask, rgba[:, 3:])
    assert _split_hint(torch.cat([rgba[:, :3], torch.ones(1, 1, 8, 8)], 1), None)[1] is None
    assert _split_hint(rgba[:, :1], None)[0].shape[1] == 3


@pytest.mark.parametrize("impl", ["product", "oracle"])
def test_unet_with_t2i_style_states_match_reference(impl):
    """Style adapters: context tokens appended to the guided side, the unconditional side padded with its own tail
    (core.py:221-237), next to a standard adapter."""
    from fakes import FakeHintStyleAdapter
    H = _impls()[impl]
    G = torch.load(os.path.join(GOLD, "hints.pt"))
    I = G["inputs"]
    assert len(G["t2i_style"]) == 6

    class CtxUNet(FakeHintUNet):
        def __call__(self, latents, t_, **kw):
            w = torch.arange(1, kw["encoder_hidden_states"].shape[1] + 1, dtype=torch.float32)[None, :, None]
            return super().__call__(latents, t_, **kw) + (kw["encoder_hidden_states"] * w).mean(dim=(1, 2))[:, None, None, None]
    for key, ref in G["t2i_style"].items():
        meta, n_style = key.split("/")
        ads = [FakeHintAdapter(1000, False)] + [FakeHintStyleAdapter(50 + i, tokens=2) for i in range(int(n_style))]
        e = torch.cat([I["ehs"][:1], I["ehs"][1:2]]) if meta == "f" else I["ehs"][:1]
        l = torch.cat([I["lat"][:1], I["lat"][:1]]) if meta == "f" else I["lat"][:1]
        t = I["t"][:2] if meta == "f" else I["t"][:1]
        got = H.UNetWithT2I(CtxUNet(), ads)(l, t, encoder_hidden_states=e, cfg_meta=meta)
        assert torch.equal(got, ref), key
