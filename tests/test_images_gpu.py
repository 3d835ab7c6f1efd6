"""Outpaint image tail on the device against the reference's numpy histogram matching (tests/golden/images.pt, produced by
scripts/make_golden.py from gyre/match_histograms.py inside the statements of unified_pipeline.py:2493-2510)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "images.pt")


@pytest.mark.parametrize("name", ["small/fp16", "batch3/fp16"])
def test_outpaint_histogram_match_bit_exact(name):
    from gyre_b200.images import match_histograms_outpaint
    v = torch.load(GOLD)[name]
    out = match_histograms_outpaint(v["result"].cuda(), v["source"].cuda(), v["outmask"].cuda())
    assert out.dtype == torch.float16
    diff = (out.cpu().float() - v["final"].float()).abs()
    assert torch.equal(out.cpu(), v["final"]), f"{int((diff > 0).sum())} of {diff.numel()} pixels differ, max {diff.max().item()}"
    # broadcast inputs ([1, C, H, W] source / mask with more channels, as the pipeline holds them) take the same path
    src4 = torch.cat([v["source"][:1], torch.ones_like(v["source"][:1, :1])], dim=1)
    out2 = match_histograms_outpaint(v["result"].cuda(), src4.cuda(), v["outmask"][:1].cuda())
    assert torch.equal(out2.cpu(), v["final"])


def test_to_uint8_nhwc():
    from gyre_b200.images import to_uint8_nhwc
    x = torch.rand(2, 3, 8, 8, generator=torch.Generator().manual_seed(1)).half()
    ref = (x.permute(0, 2, 3, 1).to(torch.float32) * 255).round().to(torch.uint8)
    assert torch.equal(to_uint8_nhwc(x.cuda()).cpu(), ref)


def test_pipeline_outpaint_tail():
    """`B200Pipeline(..., image=, mask_image=, outmask_image=)`: the decoded image goes through the device histogram match +
    source composite (unified_pipeline.py:2493-2510) - equal to applying the reference's numpy statements to the image the
    same request gives without `outmask_image`, in "pt" and in "uint8" output."""
    import numpy as np
    from oracle.unet import UNetConfig, synth_params, unet_param_shapes
    from oracle.vae import VAEConfig, vae_param_shapes
    from gyre_b200.pipeline import B200Pipeline
    from gyre_b200.unet import B200UNet
    from gyre_b200.vae import B200VAE
    cfg, vcfg = UNetConfig.tiny(), VAEConfig.tiny()
    pipe = B200Pipeline(B200UNet(cfg).load_state_dict(synth_params(unet_param_shapes(cfg), seed=1234)),
                        B200VAE(vcfg).load_state_dict(synth_params(vae_param_shapes(vcfg), seed=4321)))
    pipe.unet_sample_size_override = 16
    g = torch.Generator().manual_seed(5)
    emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=g).cuda()
    unc = torch.randn(2, 77, cfg.cross_attention_dim, generator=g).cuda()
    image = torch.rand(1, 3, 128, 128, generator=g).cuda()
    mask = torch.zeros(1, 1, 128, 128).cuda()
    mask[:, :, 32:, 40:] = 1.0
    outmask = torch.zeros(1, 3, 128, 128).cuda()
    outmask[:, :, 32:, 40:] = 1.0
    outmask[:, :, 32:48, 40:56] = 0.5

    def run(**kw):
        return pipe(emb, unc, height=128, width=128, num_inference_steps=5, guidance_scale=7.5,
                    generator=[torch.Generator("cpu").manual_seed(s) for s in (7, 8)], sampler="k_euler_ancestral",
                    image=image, mask_image=mask, strength=0.8, **kw).images
    plain = run(output_type="pt")
    out = run(output_type="pt", outmask_image=outmask)
    out_u8 = run(output_type="uint8", outmask_image=outmask)
    # the reference's statements on the plain result (fp16 tensors on the CPU, numpy histogram match restated from
    # gyre/match_histograms.py:12-37 for the uint8 branch)
    res = plain.cpu()
    src = image.cpu().half().expand(2, -1, -1, -1)
    om = outmask.cpu().half().expand(2, -1, -1, -1)
    ref_img = src * (1 - om) + res * om
    q = lambda t: (t.permute(0, 2, 3, 1).to(torch.float32) * 255).round().to(torch.uint8).numpy()
    a, b = q(res), q(ref_img)
    matched = np.empty(a.shape, dtype=a.dtype)
    for ch in range(3):
        s_, t_ = a[..., ch], b[..., ch]
        sc = np.bincount(s_.reshape(-1))
        tc = np.bincount(t_.reshape(-1))
        tv = np.nonzero(tc)[0]
        tc = tc[tv]
        matched[..., ch] = np.interp(np.cumsum(sc) / s_.size, np.cumsum(tc) / t_.size, tv)[s_.reshape(-1)].reshape(s_.shape)
    m = (torch.from_numpy(matched).to(torch.float32) / 255.0).permute(0, 3, 1, 2).to(res.dtype)
    final = src * (1 - om) + m * om
    assert torch.equal(out.cpu(), final)
    assert torch.equal(out_u8.cpu(), (final.permute(0, 2, 3, 1).to(torch.float32) * 255).round().to(torch.uint8))
    assert not torch.equal(out, plain)


def test_images_resize_matches_reference_fixture():
    """gyre_b200.images.resize (lanczos3 ResizeRight on the device) against outputs of the reference's gyre/images.py:resize
    (tests/golden/resize.pt, scripts/make_golden.py:pin_resize): down / up / anisotropic scales, both sharpness modes, hard masks."""
    import os
    from gyre_b200.images import resize
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "resize.pt"))
    assert len(G) == 9
    for c in G:
        got = resize(c["x"].cuda(), c["factors"], sharpness=c["sharpness"])
        assert got.shape == c["out"].shape and got.dtype == c["x"].dtype
        err = (got.cpu() - c["out"]).abs().max().item()
        assert err < 2e-6, (tuple(c["x"].shape), c["factors"], c["sharpness"], err)
    x16 = G[0]["x"].half().cuda()
    out16 = resize(x16, G[0]["factors"])
    assert out16.dtype == torch.float16 and (out16.float().cpu() - G[0]["out"]).abs().max().item() < 2e-3
    with pytest.raises(NotImplementedError):
        resize(x16, 0.5, sharpness=0)
    with pytest.raises(ValueError):
        resize(torch.rand(1, 1, 64, 64).cuda(), 1 / 64)              # window wider than the image: the reference fails too
