"""T2I-adapter encoder on the native kernels against vectors of the oracle that scripts/make_golden.py pinned bit for bit
to the reference's own gyre/pipeline/t2i_adapter/adapter.py (tests/golden/t2i_adapter.pt), and chained into the native UNet."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "t2i_adapter.pt")


def rel_err(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


@pytest.mark.parametrize("name", ["main_tiny", "conv_tiny"])
def test_adapter_vs_reference_vectors(name):
    from gyre_b200.t2i_adapter import B200T2iAdapter, adapter_param_shapes
    v = torch.load(GOLD)[name]
    kw = v["config"]
    assert adapter_param_shapes(**kw) == {k: tuple(t.shape) for k, t in v["state_dict"].items()}
    ad = B200T2iAdapter(**kw).load_state_dict(v["state_dict"])
    feats = ad(v["x"].cuda())
    assert len(feats) == len(v["features"])
    errs = []
    for mine, ref in zip(feats, v["features"]):
        assert tuple(mine.shape) == tuple(ref.shape)
        errs.append(rel_err(mine.cpu(), ref))
    print(f"t2i adapter {name}: rel err per level {['%.2e' % e for e in errs]}")
    assert max(errs) < 4e-3          # fp16 activations between ~20 convolutions, fp32 accumulation


def test_light_adapter_vs_reference_vectors():
    """`type: light` (Adapter_light): vectors of the oracle pinned bit for bit to the reference class."""
    from gyre_b200.t2i_adapter import B200T2iAdapter
    v = torch.load(GOLD)["light_tiny"]
    ad = B200T2iAdapter.light(**v["config"]).load_state_dict(v["state_dict"])
    feats = ad(v["x"].cuda())
    assert [tuple(f.shape) for f in feats] == [tuple(r.shape) for r in v["features"]]
    errs = [rel_err(mine.cpu(), ref) for mine, ref in zip(feats, v["features"])]
    print(f"light t2i adapter: rel err per level {['%.2e' % e for e in errs]}")
    assert max(errs) < 4e-3
    with pytest.raises(Exception):
        B200T2iAdapter.light(channels=(40, 80, 120, 120))          # channels / 4 must stay a multiple of 8


def test_adapter_into_unet_vs_oracle():
    """Native adapter -> native UNet (`adapter_states=`) == oracle adapter -> oracle UNet, SD-shaped widths in miniature."""
    from oracle import t2i_adapter as oad
    from oracle.unet import UNetConfig, synth_params, unet_forward, unet_param_shapes
    from gyre_b200.t2i_adapter import B200T2iAdapter
    from gyre_b200.unet import B200UNet
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = UNetConfig.tiny()
    PU = synth_params(unet_param_shapes(cfg), seed=1234)
    kw = dict(channels=list(cfg.block_out_channels), nums_rb=2, cin=192, ksize=1, sk=True, use_conv=False)
    PA = synth_params(oad.adapter_param_shapes(**kw), seed=55)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 4, 16, 16, generator=g).half()
    ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=g).half()
    hint = torch.rand(2, 3, 128, 128, generator=g).half()
    t = torch.tensor([700, 300])
    with torch.no_grad():
        states = oad.adapter_forward(PA, hint.float(), **{k: v for k, v in kw.items() if k != "cin"})
        ref = unet_forward(PU, cfg, x.float(), t, ctx.float(), adapter_states=states)
        plain = unet_forward(PU, cfg, x.float(), t, ctx.float())
    ad = B200T2iAdapter(**kw).load_state_dict(PA)
    unet = B200UNet(cfg).load_state_dict(PU)
    out = unet(x.cuda(), t.cuda(), encoder_hidden_states=ctx.cuda(), adapter_states=ad(hint.cuda())).sample
    err, moved = rel_err(out.cpu(), ref), rel_err(ref, plain)
    print(f"t2i adapter -> unet: rel err {err:.3e}; the adapter moves the output by {moved:.3e}")
    assert err < 5e-3 and moved > 5e-2


def test_adapter_autoinvert_and_errors():
    from gyre_b200.t2i_adapter import B200T2iAdapter
    v = torch.load(GOLD)["main_tiny"]
    kw = v["config"]
    plain = B200T2iAdapter(**kw).load_state_dict(v["state_dict"])
    inv = B200T2iAdapter(autoinvert=True, **kw).load_state_dict(v["state_dict"])
    white = (0.9 + 0.1 * v["x"]).cuda()                 # mostly white: inverted before the encoder
    a, b = inv(white), plain(1 - white)
    assert all(torch.equal(p, q) for p, q in zip(a, b))
    dark = (0.3 * v["x"]).cuda()
    assert all(torch.equal(p, q) for p, q in zip(inv(dark), plain(dark)))
    with pytest.raises(ValueError):
        plain(torch.zeros(1, 1, 64, 64).half().cuda())
    with pytest.raises(ValueError):
        plain(torch.zeros(1, 3, 60, 64).half().cuda())


def test_style_adapter_vs_reference_vectors():
    """`type: style` (StyleAdapter): tokens of the oracle pinned to the reference class (real nn.MultiheadAttention)."""
    from gyre_b200.clip_vision import B200T2iStyleAdapter, style_adapter_param_shapes
    v = torch.load(GOLD)["style_tiny"]
    assert style_adapter_param_shapes(**v["config"]) == {k: tuple(t.shape) for k, t in v["state_dict"].items()}
    ad = B200T2iStyleAdapter(**v["config"]).load_state_dict(v["state_dict"])
    tok = ad(v["x"].cuda())
    assert tuple(tok.shape) == tuple(v["tokens"].shape)
    err = rel_err(tok.cpu(), v["tokens"])
    print(f"style adapter: rel err {err:.2e}")
    assert err < 5e-3
    # one sample's tokens do not depend on its batch neighbour
    assert torch.equal(ad(v["x"][1:].cuda()), tok[1:])


def test_clip_vision_hidden_states_vs_transformers_fixture():
    """B200CLIPVisionModel (`clip_model.vision_model(image, output_hidden_states=)`): last and penultimate hidden states against
    the oracle pinned to transformers' CLIPVisionModel (tests/golden/safety.pt)."""
    from gyre_b200.clip_vision import B200CLIPVisionModel
    m = torch.load(os.path.join(os.path.dirname(__file__), "golden", "safety.pt"))["models"]["tiny"]
    sd = {k[len("vision_model."):]: v for k, v in m["state_dict"].items() if k.startswith("vision_model.vision_model.")}
    vm = B200CLIPVisionModel(m["vision_config"]).load_state_dict(sd)
    out = vm.vision_model(m["clip_input"][:2].cuda(), output_hidden_states=True, return_dict=True)
    for got, ref in ((out.last_hidden_state, m["hidden_last"]), (out.hidden_states[-1], m["hidden_last"]),
                     (out.hidden_states[-2], m["hidden_penultimate"])):
        err = (got.float().cpu() - ref.float()).abs().max().item()
        assert err < 2e-2 * max(1.0, ref.float().abs().max().item()), err
    assert len(out.hidden_states) == m["vision_config"]["num_hidden_layers"] + 1
    with pytest.raises(Exception):
        from gyre_b200 import _native as N
        import ctypes as C
        N.check(N.load().gyre_b200_safety_scores(vm._h, None, 1, None, None, None, 0, None), "safety_scores")   # tower-only handle
