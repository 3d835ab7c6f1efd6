"""The oracle's restatement of the host-side hot segment against vectors produced by the REFERENCE'S OWN CODE
(scripts/make_golden.py pin_segment / pin_hint_classes / pin_call ran gyre/pipeline/unified_pipeline.py -
UnifiedPipeline.__call__, the mode classes, the mode tree with the hires fix and graft, UnifiedPipelineHint_* - and
common_scheduler.py - KDiffusionScheduler / DiffusersScheduler - from /root/reference over the oracle UNet / VAE, and asserted
agreement there).  These tests re-run the oracle against the committed results so that it cannot drift."""
import os

import pytest
import torch

from oracle import hires as ohires
from oracle import sampling as osamp
from oracle.unet import OracleUNet, UNetConfig, synth_params, unet_param_shapes
from oracle.vae import OracleVAE, VAEConfig, vae_param_shapes

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SEEDS = [420420420, 420420421]


@pytest.fixture(scope="module")
def models():
    cfg = UNetConfig.tiny()
    unet = OracleUNet(cfg, synth_params(unet_param_shapes(cfg), seed=1234))
    cfg9 = UNetConfig.tiny(in_channels=9)
    unet9 = OracleUNet(cfg9, synth_params(unet_param_shapes(cfg9), seed=1234))
    vcfg = VAEConfig.tiny()
    VP = synth_params(vae_param_shapes(vcfg), seed=4321)
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    unc = torch.randn(1, 77, cfg.cross_attention_dim, generator=torch.Generator().manual_seed(12)).expand(2, -1, -1).contiguous()
    return dict(cfg=cfg, unet=unet, unet9=unet9, vae=lambda: OracleVAE(vcfg, VP), emb=emb, unc=unc)


def _rel(a, b):
    return (a.float() - b.float()).abs().max().item() / b.float().abs().max().item()


SEGMENT_KEYS = ["txt2img/euler_a/6/128x128/parallel/float32", "txt2img/euler_a/5/128x128/sequential/float32",
                "txt2img/euler/5/192x128/parallel/float32", "txt2img/euler_a/4/64x128/parallel/float32",
                "txt2img/heun/4/128x128/parallel/float32", "txt2img/dpm_2_a/4/128x128/parallel/float32",
                "txt2img/lms/5/128x128/parallel/float32", "txt2img/dpmpp_2s_a/4/128x128/parallel/float32",
                "txt2img/dpmpp_2m/5/128x128/parallel/float32"]


@pytest.mark.parametrize("key", SEGMENT_KEYS)
def test_txt2img_segment_matches_reference_classes(models, key):
    G = torch.load(os.path.join(GOLD, "segment.pt"))
    _, name, steps, hw, execution, _ = key.split("/")
    H, W = (int(v) for v in hw.split("x"))
    cfgu = (osamp.CFGParallel if execution == "parallel" else osamp.CFGSequential)(models["unet"], models["unc"], models["emb"], 7.5)
    with torch.no_grad():
        mine = osamp.txt2img_latents(cfgu, batch=2, in_channels=4, height=H, width=W, sample_size=16, seeds=SEEDS,
                                     steps=int(steps), sampler=name)
    assert _rel(mine, G[key]) < 2e-5, key


@pytest.mark.parametrize("kind,strength", [("img2img", 0.6), ("runway", 0.75), ("runway", 1.0), ("inpaint", 0.8)])
def test_image_modes_match_reference_classes(models, kind, strength):
    G = torch.load(os.path.join(GOLD, "segment.pt"))
    cfg = models["cfg"]
    gi = torch.Generator().manual_seed(21)
    emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=gi)
    unc = torch.randn(1, 77, cfg.cross_attention_dim, generator=gi).expand(2, -1, -1).contiguous()
    image = torch.rand(1, 3, 128, 128, generator=gi)
    mask = torch.zeros(1, 1, 128, 128)
    mask[:, :, 32:96, 40:104] = 1.0
    u = models["unet9"] if kind == "runway" else models["unet"]
    kw = {} if kind == "img2img" else {"mask_image": mask}
    with torch.no_grad():
        mine = osamp.image_mode_latents(u, models["vae"](), unc, emb, 7.5, image=image, seeds=SEEDS, steps=10, strength=strength, **kw)
    assert _rel(mine, G[f"{kind}/{strength}"]) < 2e-5


def test_whole_call_compositions_match_reference_call(models):
    """UnifiedPipeline.__call__ itself: txt2img, the hires fix (on by default above the native size) and its off switch,
    grafted inpaint, hires over img2img / Runway inpaint, DDIM through the DiffusersScheduler wrapper."""
    G = torch.load(os.path.join(GOLD, "call.pt"))
    m = models
    cfgu = osamp.CFGParallel(m["unet"], m["unc"], m["emb"], 7.5)
    gi = torch.Generator().manual_seed(21)
    image = torch.rand(1, 3, 128, 128, generator=gi)
    mask = torch.zeros(1, 1, 128, 128)
    mask[:, :, 32:96, 40:104] = 1.0
    big_image = torch.rand(1, 3, 192, 192, generator=gi)
    big_mask = torch.zeros(1, 1, 192, 192)
    big_mask[:, :, 48:144, 60:156] = 1.0
    with torch.no_grad():
        runs = {
            "hires txt2img 256x192": lambda: ohires.hires_txt2img_latents(cfgu, batch=2, height=256, width=192, sample_size=16,
                                                                          seeds=SEEDS, steps=5, oos_fraction=0.6),
            "txt2img 192 hires off": lambda: osamp.txt2img_latents(cfgu, batch=2, in_channels=4, height=192, width=192,
                                                                   sample_size=16, seeds=SEEDS, steps=5, sampler="euler_a"),
            "grafted inpaint": lambda: ohires.grafted_inpaint_latents(m["unet9"], m["unet"], m["vae"](), m["unc"], m["emb"], 7.5,
                                                                      image=image, mask_image=mask, seeds=SEEDS, steps=8,
                                                                      strength=0.75),
            "hires runway inpaint": lambda: ohires.hires_image_mode_latents(m["unet9"], m["vae"](), m["unc"], m["emb"], 7.5,
                                                                            image=big_image, mask_image=big_mask, seeds=SEEDS,
                                                                            steps=6, strength=0.75, sample_size=16,
                                                                            oos_fraction=1.0),
            "ddim": lambda: osamp.txt2img_latents(cfgu, batch=2, in_channels=4, height=128, width=128, sample_size=16,
                                                  seeds=SEEDS, steps=8, sampler="ddim"),
        }
        cfg5 = UNetConfig.tiny(in_channels=5)
        unet5 = OracleUNet(cfg5, synth_params(unet_param_shapes(cfg5), seed=321))
        blend = {"start": 0.15, "end": 0.75, "easing": "sine"}
        # the depth map is what the reference's own lines made of the hint image (2 * images.resize(.., 1/8, sharpness=2) - 1)
        runs["grafted depth"] = lambda: ohires.depth_txt2img_latents(unet5, m["unet"], m["unc"], m["emb"], 7.5,
                                                                     depth_map=G["depth_map"], seeds=SEEDS, steps=7, sample_size=16,
                                                                     height=128, width=128, graft_blend=blend)
        runs["depth"] = lambda: ohires.depth_txt2img_latents(unet5, m["unet"], m["unc"], m["emb"], 7.5, depth_map=G["depth_map"],
                                                             seeds=SEEDS, steps=7, sample_size=16, height=128, width=128)
        assert set(runs) <= set(G) and len(G) == 13 + 2
        for key, fn in runs.items():
            assert _rel(fn(), G[key]) < 2e-6, key


def test_hint_classes_match_reference_classes(models):
    from types import SimpleNamespace as SN
    from oracle import controlnet as ocn
    from oracle import hints as oh
    from oracle import t2i_adapter as oad
    G = torch.load(os.path.join(GOLD, "hint_classes.pt"))
    cfg = models["cfg"]
    Pcn = synth_params(ocn.controlnet_param_shapes(cfg), seed=77)
    akw = dict(channels=list(cfg.block_out_channels), nums_rb=2, cin=192, ksize=1, sk=True, use_conv=False)
    Pad = synth_params(oad.adapter_param_shapes(**akw), seed=91)
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    unc = torch.randn(1, 77, cfg.cross_attention_dim, generator=g).expand(2, -1, -1).contiguous()
    img = torch.rand(1, 3, 128, 128, generator=g).half().float()

    def cn(cnlatents, t, encoder_hidden_states, controlnet_cond):
        down, mid = ocn.controlnet_forward(Pcn, cfg, cnlatents, t, encoder_hidden_states, controlnet_cond)
        return SN(down_block_res_samples=down, mid_block_res_sample=mid)

    def ad(x):
        return oad.adapter_forward(Pad, x, **{k: v for k, v in akw.items() if k != "cin"})
    cases = {"controlnet soft 0.7": ([oh.ControlnetHint(cn, img, weight=0.7, soft_injection=True, cfg_only=False)], True),
             "controlnet cfg_only sequential": ([oh.ControlnetHint(cn, img, weight=0.8, soft_injection=False, cfg_only=True)], False),
             "t2i soft cfg_only + controlnet": ([oh.ControlnetHint(cn, img, weight=0.5, soft_injection=False, cfg_only=False),
                                                 oh.T2iHint(ad, img.expand(2, -1, -1, -1), weight=0.9, soft_injection=True,
                                                            cfg_only=True)], True)}
    assert set(cases) <= set(G) and len([k for k in G if not k.startswith('r') and k != 'style_image']) == 11
    for key, (hints, parallel) in cases.items():
        eps = oh.guided_eps_unet(models["unet"], unc, emb, 7.5, hints, parallel=parallel)
        with torch.no_grad():
            mine = osamp.txt2img_latents(eps, batch=2, in_channels=4, height=128, width=128, sample_size=16, seeds=SEEDS, steps=5,
                                         sampler="euler_a")
        assert _rel(mine, G[key]) < 2e-5, key


def test_masked_hints_and_controlnet_under_inpaint_match_reference_classes(models):
    """RGBA hints (mask = alpha, resized to every residual / state with images.resize) and a ControlNet under the 9-channel
    inpaint UNet, against the reference's hint classes run inside its own stack (hint_classes.pt)."""
    from types import SimpleNamespace as SN
    from oracle import controlnet as ocn
    from oracle import hints as oh
    from oracle import t2i_adapter as oad
    G = torch.load(os.path.join(GOLD, "hint_classes.pt"))
    cfg = models["cfg"]
    Pcn = synth_params(ocn.controlnet_param_shapes(cfg), seed=77)
    akw = dict(channels=list(cfg.block_out_channels), nums_rb=2, cin=192, ksize=1, sk=True, use_conv=False)
    Pad = synth_params(oad.adapter_param_shapes(**akw), seed=91)
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    unc = torch.randn(1, 77, cfg.cross_attention_dim, generator=g).expand(2, -1, -1).contiguous()
    rgba = G["rgba"].float()

    def cn(cnlatents, t, encoder_hidden_states, controlnet_cond):
        down, mid = ocn.controlnet_forward(Pcn, cfg, cnlatents, t, encoder_hidden_states, controlnet_cond)
        return SN(down_block_res_samples=down, mid_block_res_sample=mid)

    def ad(x):
        return oad.adapter_forward(Pad, x, **{k: v for k, v in akw.items() if k != "cin"})
    with torch.no_grad():
        eps = oh.guided_eps_unet(models["unet"], unc, emb, 7.5,
                                 [oh.ControlnetHint(cn, rgba, weight=0.9, soft_injection=True, cfg_only=False),
                                  oh.T2iHint(ad, rgba.expand(2, -1, -1, -1), weight=0.8)])
        mine = osamp.txt2img_latents(eps, batch=2, in_channels=4, height=256, width=256, sample_size=16, seeds=SEEDS, steps=4,
                                     sampler="euler_a")
        assert _rel(mine, G["masked controlnet + t2i"]) < 2e-5
        mine = osamp.image_mode_latents(models["unet9"], models["vae"](), unc, emb, 7.5, image=G["runway_image"].float(),
                                        mask_image=G["runway_mask"].float(), seeds=SEEDS, steps=4, strength=0.8,
                                        hints=[oh.ControlnetHint(cn, rgba, weight=1.0)])
        assert _rel(mine, G["controlnet under runway inpaint"]) < 2e-5


def test_images_resize_oracle_matches_reference_fixture():
    G = torch.load(os.path.join(GOLD, "resize.pt"))
    for c in G:
        out = ohires.images_resize(c["x"], c["factors"], sharpness=c["sharpness"])
        assert out.shape == c["out"].shape and (out - c["out"]).abs().max().item() <= 5e-7
