"""gyre_b200.safety_checker (B200FeatureExtractor + B200SafetyChecker) against the fixtures pinned to Pillow and to the
reference's FlagOnlySafetyChecker (tests/golden/safety.pt, scripts/make_golden.py:pin_safety), against the oracle on seeded
inputs, and at ViT-L/14 size through properties."""
import os

import numpy as np
import pytest
import torch

from oracle import safety as osf

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "safety.pt"))


def test_feature_extractor_bit_exact_against_pillow_fixture(gold):
    from gyre_b200.safety_checker import B200FeatureExtractor
    fx = B200FeatureExtractor()
    for r in gold["resize"]:
        h, w = r["image_hw"]
        img = torch.from_numpy(osf.synthetic_image(h, w)).cuda()
        nh, nw = fx.output_size(h, w)
        assert (nh, nw) == tuple(r["size"])
        resized = fx.resize(img[None], nh, nw)[0].cpu()
        assert int(resized.long().sum()) == r["resized_sum"], (h, w)
        if r["resized"] is not None:
            assert torch.equal(resized, r["resized"])
        pv = fx(img[None]).pixel_values
        assert pv.dtype == torch.float16 and tuple(pv.shape) == (1, 3, 224, 224)
        assert torch.equal(pv[0].cpu(), r["pixel_values_f16"]), (h, w)


def test_feature_extractor_batch_and_float_input_match_oracle():
    from gyre_b200.safety_checker import B200FeatureExtractor
    fx = B200FeatureExtractor()
    g = torch.Generator().manual_seed(4)
    x = torch.rand(3, 3, 192, 320, generator=g)                               # decoded images in [0, 1]
    u8 = (x.permute(0, 2, 3, 1) * 255).round().to(torch.uint8)                # numpy_to_pil's quantisation
    ref = torch.from_numpy(osf.clip_preprocess(u8.numpy())).half()
    got = fx(x.cuda(), return_tensors="pt").to("cuda").pixel_values
    assert torch.equal(got.cpu(), ref)
    assert torch.equal(fx(u8.cuda()).pixel_values.cpu(), ref)
    with pytest.raises(ValueError):
        fx(u8[..., :2].cuda())


def _checker(m):
    from gyre_b200.safety_checker import B200SafetyChecker
    cfg = {"vision_config": m["vision_config"], "projection_dim": m["projection_dim"]}
    return B200SafetyChecker(cfg).load_state_dict(m["state_dict"])


@pytest.mark.parametrize("name", ["tiny", "vit224", "gelu"])
def test_safety_checker_matches_reference_fixture(gold, name):
    m = gold["models"][name]
    sc = _checker(m)
    scores, embeds = sc.scores(m["clip_input"].cuda(), return_embeds=True)
    ref_e = m["image_embeds"]
    rel = (embeds.float().cpu() - ref_e).norm() / ref_e.norm()
    assert rel < 5e-3, f"image_embeds rel err {rel}"
    err = (scores.cpu() - m["scores"]).abs().max().item()
    assert err < 1e-3, f"cosine scores: max abs err {err}"    # fp16 tower vs the fp32 reference (measured 1.9 - 3.3e-4); margins >= 8e-3
    images = np.zeros((scores.shape[0], 8, 8, 3), np.float32)
    out, flags = sc(clip_input=m["clip_input"].cuda(), images=images)
    assert out is images and flags == m["flags"]
    for r, g in zip(sc.last_result, m["result"]):
        assert r["bad_concepts"] == g["bad_concepts"]
        assert np.abs(np.array(list(r["concept_scores"].values())) - np.array(g["concept_scores"])).max() < 2e-3
        assert np.abs(np.array(list(r["special_scores"].values())) - np.array(g["special_scores"])).max() < 2e-3


def test_safety_checker_errors(gold):
    from gyre_b200 import _native as N
    from gyre_b200.safety_checker import B200SafetyChecker
    m = gold["models"]["tiny"]
    cfg = {"vision_config": m["vision_config"], "projection_dim": m["projection_dim"]}
    sc = B200SafetyChecker(cfg)
    with pytest.raises(N.NativeError):
        sc.scores(m["clip_input"].cuda())                                     # weights not loaded
    partial = {k: v for k, v in m["state_dict"].items() if "fc2" not in k}
    with pytest.raises(N.NativeError):
        B200SafetyChecker(cfg).load_state_dict(partial)                       # finalize: missing parameters
    sc.load_state_dict(m["state_dict"])
    with pytest.raises(ValueError):
        sc.scores(torch.zeros(1, 3, 32, 32).cuda())
    with pytest.raises(Exception):
        B200SafetyChecker({"vision_config": dict(m["vision_config"], hidden_act="relu"), "projection_dim": 32})


def test_safety_checker_vit_l14_properties():
    """Full CompVis checker size (ViT-L/14, 257 tokens, 24 layers): against the oracle on a small batch, and batch
    composition must not change an image's scores."""
    from gyre_b200.safety_checker import B200SafetyChecker, ClipVisionConfig, safety_checker_param_shapes
    cfg = ClipVisionConfig.vit_l14()
    g = torch.Generator().manual_seed(31)
    sd = {}
    for k, shp in safety_checker_param_shapes(cfg).items():
        if k.endswith("norm.weight") or "layer_norm" in k and k.endswith("weight") or k.endswith("layrnorm.weight"):
            sd[k] = 1 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.05 * torch.randn(shp, generator=g)
        elif k.endswith("embeds"):
            sd[k] = torch.randn(shp, generator=g)
        elif k.endswith("embeds_weights"):
            sd[k] = torch.full(shp, 0.1)
        elif k.endswith("patch_embedding.weight"):
            sd[k] = (0.05 * torch.randn(shp, generator=g)).half().float()
        elif k.endswith(("class_embedding", "position_embedding.weight")):
            sd[k] = (0.3 * torch.randn(shp, generator=g)).half().float()
        else:
            sd[k] = (torch.randn(shp, generator=g) * (1.5 if k.endswith(("v_proj.weight", "out_proj.weight", "fc2.weight")) else 1.0)
                     / shp[-1] ** 0.5).half().float()
    sc = B200SafetyChecker(cfg).load_state_dict(sd)
    x = (torch.randn(5, 3, 224, 224, generator=g) * torch.linspace(0.5, 2.0, 5)[:, None, None, None]).half()
    scores, emb = sc.scores(x.cuda(), return_embeds=True)
    assert torch.isfinite(scores).all() and scores.abs().max() <= 1.0 + 1e-3
    P = {k[len("vision_model."):] if k.startswith("vision_model.vision_model.") else k: v for k, v in sd.items()}
    with torch.no_grad():
        _, emb_ref = osf.clip_vision_forward(P, x[:2].float(), num_layers=24, num_heads=16, patch_size=14)
        ref = osf.cosine_scores(emb_ref, P)
    err = (scores[:2].cpu() - ref).abs().max().item()
    assert err < 1e-2, f"ViT-L/14 cosine scores vs oracle: {err}"
    # an image's scores do not depend on its neighbours in the batch or its position
    perm = torch.tensor([3, 0, 4, 1, 2])
    s2 = sc.scores(x[perm].cuda())
    assert torch.equal(s2.cpu(), scores[perm].cpu())
    s1 = sc.scores(x[2:3].cuda())
    assert (s1.cpu() - scores[2:3].cpu()).abs().max().item() < 2e-3         # (tile scheduling may differ with M)


def test_pipeline_reports_nsfw_flags(gold):
    """The pipeline tail: decode -> 8-bit quantisation -> feature extractor -> checker, all on the device, equals the oracle
    chain on the decoded image."""
    from oracle.unet import UNetConfig, synth_params, unet_param_shapes
    from oracle.vae import VAEConfig, vae_param_shapes
    from gyre_b200.pipeline import B200Pipeline
    from gyre_b200.safety_checker import B200FeatureExtractor
    from gyre_b200.unet import B200UNet
    from gyre_b200.vae import B200VAE
    ucfg, vcfg = UNetConfig.tiny(), VAEConfig.tiny()
    unet = B200UNet(ucfg).load_state_dict(synth_params(unet_param_shapes(ucfg), seed=1234))
    vae = B200VAE(vcfg).load_state_dict(synth_params(vae_param_shapes(vcfg), seed=4321))
    m = gold["models"]["tiny"]
    sc = _checker(m)
    pipe = B200Pipeline(unet, vae, safety_checker=sc,
                        feature_extractor=B200FeatureExtractor(size=m["vision_config"]["image_size"]))
    pipe.unet_sample_size_override = 16
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 77, ucfg.cross_attention_dim, generator=g).half().cuda()
    unc = torch.randn(2, 77, ucfg.cross_attention_dim, generator=g).half().cuda()
    gens = [torch.Generator("cpu").manual_seed(s) for s in (5, 6)]
    out = pipe(emb, unc, height=128, width=128, num_inference_steps=4, generator=gens, sampler="k_euler")
    assert out.nsfw_content_detected is not None and len(out.nsfw_content_detected) == 2
    u8 = (out.images.float().permute(0, 2, 3, 1) * 255).round().to(torch.uint8).cpu().numpy()
    S = m["vision_config"]["image_size"]
    pv = torch.from_numpy(osf.clip_preprocess(u8, size=S)).half()
    ref_scores = sc.scores(pv.cuda())
    res, flags = osf.flag_only(ref_scores.cpu().numpy(), m["state_dict"]["special_care_embeds_weights"].float(),
                               m["state_dict"]["concept_embeds_weights"].float())
    assert out.nsfw_content_detected == flags
    assert [r["bad_concepts"] for r in sc.last_result] == [r["bad_concepts"] for r in res]
    gens = [torch.Generator("cpu").manual_seed(s) for s in (5, 6)]
    off = pipe(emb, unc, height=128, width=128, num_inference_steps=4, generator=gens, sampler="k_euler", run_safety_checker=False)
    assert off.nsfw_content_detected == [False, False] and torch.equal(off.images, out.images)
