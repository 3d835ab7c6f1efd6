"""ControlNet / T2I-adapter hints end to end: B200Pipeline(hints=[...]) - the native ControlNet at every UNet call, adapter
states once per request, both injected into the native UNet - against the oracle's composition of the wrapper stack
(oracle/hints.py, pinned to gyre/pipeline/unet/core.py) over the oracle ControlNet / adapter / UNet on the CPU."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from oracle import controlnet as ocn
    from oracle import t2i_adapter as oad
    from oracle.unet import OracleUNet, UNetConfig, synth_params, unet_param_shapes
    from gyre_b200.controlnet import B200ControlNet
    from gyre_b200.pipeline import B200Pipeline
    from gyre_b200.t2i_adapter import B200T2iAdapter
    from gyre_b200.unet import B200UNet
    cfg = UNetConfig.tiny()
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    Pcn = synth_params(ocn.controlnet_param_shapes(cfg), seed=77)
    akw = dict(channels=list(cfg.block_out_channels), nums_rb=2, cin=192, ksize=1, sk=True, use_conv=False)
    Pad = synth_params(oad.adapter_param_shapes(**akw), seed=91)
    pipe = B200Pipeline(B200UNet(cfg).load_state_dict(P), None)
    pipe.unet_sample_size_override = 16
    cn = B200ControlNet(cfg).load_state_dict(Pcn)
    ad = B200T2iAdapter(**akw).load_state_dict(Pad)
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    unc = torch.randn(1, 77, cfg.cross_attention_dim, generator=g).expand(2, -1, -1).contiguous()
    hint_img = torch.rand(1, 3, 128, 128, generator=g).half().float()

    def o_cn(cnlatents, t, encoder_hidden_states, controlnet_cond):
        with torch.no_grad():
            down, mid = ocn.controlnet_forward(Pcn, cfg, cnlatents, t, encoder_hidden_states, controlnet_cond)
        return SimpleNamespace(down_block_res_samples=down, mid_block_res_sample=mid)

    def o_ad(x):
        with torch.no_grad():
            return oad.adapter_forward(Pad, x, **{k: v for k, v in akw.items() if k != "cin"})
    return SimpleNamespace(cfg=cfg, pipe=pipe, cn=cn, ad=ad, o_cn=o_cn, o_ad=o_ad, o_unet=OracleUNet(cfg, P), emb=emb, unc=unc,
                           img=hint_img)


CASES = [("controlnet", dict(weight=1.0, soft_injection=False, cfg_only=False), None, "parallel"),
         ("controlnet soft 0.7", dict(weight=0.7, soft_injection=True, cfg_only=False), None, "parallel"),
         ("controlnet cfg_only", dict(weight=0.8, soft_injection=True, cfg_only=True), None, "parallel"),
         ("controlnet cfg_only sequential", dict(weight=0.8, soft_injection=False, cfg_only=True), None, "sequential"),
         ("t2i", None, dict(weight=1.0, soft_injection=False, cfg_only=False), "parallel"),
         ("t2i soft cfg_only + controlnet", dict(weight=0.5, soft_injection=False, cfg_only=False),
          dict(weight=0.9, soft_injection=True, cfg_only=True), "parallel"),
         ("t2i cfg_only sequential", None, dict(weight=1.0, soft_injection=False, cfg_only=True), "sequential")]


@pytest.mark.parametrize("name,cn_kw,ad_kw,execution", CASES, ids=[c[0] for c in CASES])
def test_pipeline_with_hints_vs_oracle(setup, name, cn_kw, ad_kw, execution):
    from oracle import hints as oh
    from oracle import sampling as osamp
    from gyre_b200.hints import B200ControlnetHint, B200T2iHint
    s = setup
    hints, ohints = [], []
    if cn_kw is not None:
        hints.append(B200ControlnetHint(s.cn, s.img.cuda(), **cn_kw))
        ohints.append(oh.ControlnetHint(s.o_cn, s.img, **cn_kw))
    if ad_kw is not None:
        hints.append(B200T2iHint(s.ad, s.img.cuda(), **ad_kw))
        # (the reference adds the states to the UNet's hidden states as they are: its hint batch has to equal the sample
        # batch; the product expands a single hint image to the batch itself)
        ohints.append(oh.T2iHint(s.o_ad, s.img.expand(2, -1, -1, -1), **ad_kw))
    seeds, steps = [420420420, 420420421], 8
    out = s.pipe(s.emb.cuda(), s.unc.cuda(), height=128, width=128, num_inference_steps=steps, guidance_scale=7.5,
                 generator=[torch.Generator("cpu").manual_seed(x) for x in seeds], sampler="k_euler_ancestral",
                 output_type="latent", latents_dtype=torch.float32, return_fp32_latents=True, hints=hints,
                 cfg_execution=execution)
    eps = oh.guided_eps_unet(s.o_unet, s.unc, s.emb, 7.5, ohints, parallel=execution == "parallel")
    with torch.no_grad():
        ref = osamp.txt2img_latents(eps, batch=2, in_channels=4, height=128, width=128, sample_size=16, seeds=seeds,
                                    steps=steps, sampler="euler_a")
        plain = osamp.txt2img_latents(oh.guided_eps_unet(s.o_unet, s.unc, s.emb, 7.5, []), batch=2, in_channels=4, height=128,
                                      width=128, sample_size=16, seeds=seeds, steps=steps, sampler="euler_a")
    got = out.latents.float().cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item() / scale
    moved = (plain - ref).abs().max().item() / scale
    assert moved > 0.05, f"the hint did not change the result ({moved}): the test would prove nothing"
    assert err < 1.4e-2, f"{name}: rel err {err} (hint moves the latents by {moved})"


def test_hints_leave_no_residuals_bound(setup):
    """A hinted run followed by a plain one equals a plain run (the residual / state pointers are per call)."""
    from gyre_b200.hints import B200ControlnetHint
    s = setup
    kw = dict(height=128, width=128, num_inference_steps=4, guidance_scale=7.5, sampler="k_euler", output_type="latent")
    gen = lambda: [torch.Generator("cpu").manual_seed(x) for x in (1, 2)]
    a = s.pipe(s.emb.cuda(), s.unc.cuda(), generator=gen(), **kw).latents
    s.pipe(s.emb.cuda(), s.unc.cuda(), generator=gen(), hints=[B200ControlnetHint(s.cn, s.img.cuda())], **kw)
    b = s.pipe(s.emb.cuda(), s.unc.cuda(), generator=gen(), **kw).latents
    assert torch.equal(a, b)


def test_hints_with_hires_fix_run(setup):
    """Above the native size the natural-size twin gets the hint image scaled like the init image (unified_pipeline.py:2148-2158)."""
    from gyre_b200.hints import B200ControlnetHint, B200T2iHint
    s = setup
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 3, 192, 192, generator=g)
    out = s.pipe(s.emb.cuda(), s.unc.cuda(), height=192, width=192, num_inference_steps=4, guidance_scale=7.5,
                 generator=[torch.Generator("cpu").manual_seed(x) for x in (1, 2)], sampler="k_euler", output_type="latent",
                 hints=[B200ControlnetHint(s.cn, img.cuda(), weight=0.5), B200T2iHint(s.ad, img.cuda())], hires_fix=True)
    assert tuple(out.latents.shape) == (2, 4, 24, 24) and torch.isfinite(out.latents.float()).all()


def test_masked_hints_vs_oracle(setup):
    """RGBA hints: the alpha channel masks every ControlNet residual / adapter state at its own resolution (images.resize,
    antialiased lanczos3, cached per request).  Oracle pinned against the reference's hint classes (hint_classes.pt)."""
    from oracle import hints as oh
    from oracle import sampling as osamp
    from gyre_b200.hints import B200ControlnetHint, B200T2iHint
    s = setup
    g = torch.Generator().manual_seed(3)
    big = torch.rand(1, 3, 256, 256, generator=g).half().float()
    alpha = torch.zeros(1, 1, 256, 256)
    alpha[:, :, 64:224, 32:160] = 1.0
    rgba = torch.cat([big, alpha], dim=1)
    seeds, steps = [420420420, 420420421], 4
    hints = [B200ControlnetHint(s.cn, rgba.cuda(), weight=0.9, soft_injection=True), B200T2iHint(s.ad, rgba.cuda(), weight=0.8)]
    assert hints[0].mask is not None and hints[1].mask is not None
    out = s.pipe(s.emb.cuda(), s.unc.cuda(), height=256, width=256, num_inference_steps=steps, guidance_scale=7.5,
                 generator=[torch.Generator("cpu").manual_seed(x) for x in seeds], sampler="k_euler_ancestral", output_type="latent",
                 latents_dtype=torch.float32, return_fp32_latents=True, hints=hints, hires_fix=False).latents
    ohints = [oh.ControlnetHint(s.o_cn, rgba, weight=0.9, soft_injection=True, cfg_only=False),
              oh.T2iHint(s.o_ad, rgba.expand(2, -1, -1, -1), weight=0.8, soft_injection=False, cfg_only=False)]
    with torch.no_grad():
        ref = osamp.txt2img_latents(oh.guided_eps_unet(s.o_unet, s.unc, s.emb, 7.5, ohints), batch=2, in_channels=4, height=256,
                                    width=256, sample_size=16, seeds=seeds, steps=steps, sampler="euler_a")
        unmasked = osamp.txt2img_latents(oh.guided_eps_unet(s.o_unet, s.unc, s.emb, 7.5, [
            oh.ControlnetHint(s.o_cn, big, weight=0.9, soft_injection=True, cfg_only=False),
            oh.T2iHint(s.o_ad, big.expand(2, -1, -1, -1), weight=0.8)]), batch=2, in_channels=4, height=256, width=256,
            sample_size=16, seeds=seeds, steps=steps, sampler="euler_a")
    scale = ref.abs().max().item()
    err = (out.float().cpu() - ref).abs().max().item() / scale
    assert (unmasked - ref).abs().max().item() / scale > 0.02, "the mask did not change the result"
    assert err < 1.4e-2, f"masked hints: rel err {err}"


def test_controlnet_under_inpaint_unet_vs_oracle(setup):
    """A ControlNet hint with the 9-channel (Runway) inpaint UNet: latents and conditioning image go through the inpaint mask
    (channel 4 of the UNet input, scaled up 8x with images.resize), residuals through its scaled-down copies."""
    from oracle import hints as oh
    from oracle import sampling as osamp
    from oracle.unet import OracleUNet, UNetConfig, synth_params, unet_param_shapes
    from oracle.vae import OracleVAE, VAEConfig, vae_param_shapes
    from gyre_b200.hints import B200ControlnetHint
    from gyre_b200.pipeline import B200Pipeline
    from gyre_b200.unet import B200UNet
    from gyre_b200.vae import B200VAE
    s = setup
    cfg9 = UNetConfig.tiny(in_channels=9)
    P9 = synth_params(unet_param_shapes(cfg9), seed=1234)
    vcfg = VAEConfig.tiny()
    VP = synth_params(vae_param_shapes(vcfg), seed=4321)
    pipe = B200Pipeline(B200UNet(cfg9).load_state_dict(P9), B200VAE(vcfg).load_state_dict(VP))
    pipe.unet_sample_size_override = 16
    g = torch.Generator().manual_seed(9)
    hint = torch.rand(1, 3, 256, 256, generator=g).half().float()
    image = torch.rand(1, 3, 256, 256, generator=g).half().float()
    mask = torch.zeros(1, 1, 256, 256)
    mask[:, :, 96:192, 64:200] = 1.0
    seeds, steps = [420420420, 420420421], 4
    out = pipe(s.emb.cuda(), s.unc.cuda(), height=256, width=256, num_inference_steps=steps, guidance_scale=7.5,
               generator=[torch.Generator("cpu").manual_seed(x) for x in seeds], sampler="k_euler_ancestral", output_type="latent",
               latents_dtype=torch.float32, return_fp32_latents=True, image=image.cuda(), mask_image=mask.cuda(), strength=0.8,
               hints=[B200ControlnetHint(s.cn, hint.cuda())], hires_fix=False).latents
    with torch.no_grad():
        kw = dict(image=image, mask_image=mask, seeds=seeds, steps=steps, strength=0.8)
        ref = osamp.image_mode_latents(OracleUNet(cfg9, P9), OracleVAE(vcfg, VP, sample_dtype=torch.float16), s.unc, s.emb, 7.5,
                                       hints=[oh.ControlnetHint(s.o_cn, hint)], **kw)
        plain = osamp.image_mode_latents(OracleUNet(cfg9, P9), OracleVAE(vcfg, VP, sample_dtype=torch.float16), s.unc, s.emb, 7.5, **kw)
    scale = ref.abs().max().item()
    err = (out.float().cpu() - ref).abs().max().item() / scale
    assert (plain - ref).abs().max().item() / scale > 0.02
    assert err < 1.4e-2, f"ControlNet under the inpaint UNet: rel err {err}"


@pytest.mark.parametrize("layer,key", [(None, "style final + t2i"), ("penultimate", "style penultimate + t2i")])
def test_style_adapter_hint_vs_reference_run(setup, layer, key):
    """A style T2I-adapter hint (rescale "cover" -> CLIP normalisation -> native CLIP vision tower -> native StyleAdapter ->
    context tokens on the guided side) next to a standard adapter, against what the REFERENCE's own UnifiedPipelineHint_T2i /
    UNetWithT2I / scheduler returned over the oracle models (tests/golden/hint_classes.pt, pin_hint_classes)."""
    import os
    from oracle import hints as oh
    from oracle import safety as osf
    from oracle import sampling as osamp
    from oracle.unet import synth_params
    from gyre_b200.clip_vision import B200CLIPVisionModel, B200T2iStyleAdapter, style_adapter_param_shapes
    from gyre_b200.hints import B200T2iHint
    s = setup
    gold = os.path.join(os.path.dirname(__file__), "golden")
    G = torch.load(os.path.join(gold, "hint_classes.pt"))
    sg = torch.load(os.path.join(gold, "safety.pt"))["models"]["tiny"]
    vis = sg["vision_config"]
    vm = B200CLIPVisionModel(vis).load_state_dict({k[len("vision_model."):]: v for k, v in sg["state_dict"].items()
                                                   if k.startswith("vision_model.vision_model.")})
    skw = dict(width=vis["hidden_size"], context_dim=s.cfg.cross_attention_dim, num_head=4, n_layes=2, num_token=4)
    st = B200T2iStyleAdapter(**skw).load_state_dict(synth_params(style_adapter_param_shapes(**skw), seed=47))
    fe = SimpleNamespace(image_mean=list(osf.CLIP_MEAN), image_std=list(osf.CLIP_STD), size={"shortest_edge": vis["image_size"]})
    # (the fixture run used these embeddings / hint image: generator 11 like the `setup` fixture)
    style_img = G["style_image"].float()
    hints = [B200T2iHint(s.ad, s.img.cuda()),
             B200T2iHint(st, style_img.cuda(), weight=0.7, cfg_only=True, clip_model=vm, feature_extractor=fe, clip_layer=layer)]
    seeds, steps = [420420420, 420420421], 4
    out = s.pipe(s.emb.cuda(), s.unc.cuda(), height=128, width=128, num_inference_steps=steps, guidance_scale=7.5,
                 generator=[torch.Generator("cpu").manual_seed(x) for x in seeds], sampler="k_euler_ancestral", output_type="latent",
                 latents_dtype=torch.float32, return_fp32_latents=True, hints=hints).latents
    ref = G[key]
    with torch.no_grad():
        no_style = osamp.txt2img_latents(oh.guided_eps_unet(s.o_unet, s.unc, s.emb, 7.5, [oh.T2iHint(s.o_ad, s.img.expand(2, -1, -1, -1))]),
                                         batch=2, in_channels=4, height=128, width=128, sample_size=16, seeds=seeds, steps=steps,
                                         sampler="euler_a")
    scale = ref.abs().max().item()
    err = (out.float().cpu() - ref).abs().max().item() / scale
    moved = (no_style - ref).abs().max().item() / scale
    assert moved > 0.01, f"the style tokens did not change the result ({moved})"
    assert err < 1.4e-2, f"style hint ({key}): rel err {err} (the style tokens move the latents by {moved})"
