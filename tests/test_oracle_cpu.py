"""The oracle against the committed golden vectors (tests/golden/*.pt).  samplers.pt / ddim.pt / tome.pt hold
outputs of the REFERENCE's own vendored code (k-diffusion, in-tree DDIM copy, ToMe), produced by
scripts/make_golden.py in the build container where /root/reference is mounted; oracle_tiny.pt pins the
oracle's UNet / VAE / pipeline outputs against accidental drift."""
import os

import pytest
import torch

from oracle import sampling as osamp
from oracle import tome as otome
from oracle.unet import UNetConfig, OracleUNet, synth_params, unet_forward, unet_param_shapes
from oracle.vae import VAEConfig, vae_decode, vae_encode_moments, vae_param_shapes

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def toy_eps(x, t):
    t = t if torch.is_tensor(t) else torch.tensor(t)
    tt = t.float().reshape(-1, *([1] * (x.ndim - 1)))
    return 0.7 * torch.tanh(x) + 0.001 * tt * x.roll(1, -1)


@pytest.mark.parametrize("name", ["euler_a", "euler", "heun", "dpmpp_2m"])
@pytest.mark.parametrize("steps", [7, 20])
@pytest.mark.parametrize("dtype_name,ldt", [("fp32", torch.float32), ("fp16", torch.float16)])
def test_samplers_match_vendored_k_diffusion(name, steps, dtype_name, ldt):
    rec = torch.load(os.path.join(GOLD, "samplers.pt"))[f"{name}/{steps}/{dtype_name}"]
    shape, seeds = rec["shape"], rec["seeds"]
    gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]
    den = osamp.EpsDenoiser(toy_eps, osamp.sd_alphas_cumprod())
    sig = osamp.k_sigmas(den, steps)
    assert torch.equal(sig, rec["sigmas"])
    x = (osamp.batched_randn(shape, gens, "cpu", ldt) * sig[0]).float()
    ns = lambda *_: osamp.batched_randn(shape, gens, "cpu", ldt).float()
    s = sig.to(ldt).float()
    if name == "euler_a":
        got = osamp.sample_euler_ancestral(den, x, s, ns)
    elif name == "euler":
        got = osamp.sample_euler(den, x, s, lambda _x: ns())
    elif name == "heun":
        got = osamp.sample_heun(den, x, s, lambda _x: ns())
    else:
        got = osamp.sample_dpmpp_2m(den, x, s, warmup_lms=True, ddim_cutoff=0.1)
    assert torch.equal(got, rec["result"]), f"{name}/{steps}/{dtype_name} drifted from the vendored k-diffusion output"


def test_denoiser_wrappers_and_schedules():
    g = torch.load(os.path.join(GOLD, "samplers.pt"))
    acp = osamp.sd_alphas_cumprod()
    v = g["vdenoiser"]
    assert torch.equal(osamp.VDenoiser(toy_eps, acp)(v["x"], v["sigma"]), v["result"])
    den = osamp.EpsDenoiser(toy_eps, acp)
    assert torch.equal(osamp.get_sigmas_karras(11, den.sigma_min, den.sigma_max, 7.0), g["karras/11"])
    st = g["sigma_to_t"]
    assert torch.equal(den.sigma_to_t(st["sigma"]), st["t"])
    # SD schedule end points
    assert abs(float(den.sigma_min) - 0.0292) < 1e-3 and abs(float(den.sigma_max) - 14.6146) < 1e-3


@pytest.mark.parametrize("pred", ["epsilon", "v_prediction"])
@pytest.mark.parametrize("eta", [0.0, 0.6])
def test_ddim_matches_in_tree_copy(pred, eta):
    rec = torch.load(os.path.join(GOLD, "ddim.pt"))[f"{pred}/{eta}"]
    gen = torch.Generator("cpu").manual_seed(99)
    got = osamp.sample_ddim(toy_eps, rec["x"].clone(), 10, osamp.sd_alphas_cumprod(), eta, gen, pred)
    assert (got - rec["result"]).abs().max().item() < 1e-6
    assert osamp.ddim_timesteps(10).tolist() == [901, 801, 701, 601, 501, 401, 301, 201, 101, 1]


def test_tome_matches_vendored_merge():
    g = torch.load(os.path.join(GOLD, "tome.pt"))
    for name, rec in g.items():
        if name == "parse_r":
            assert otome.parse_r(16, (100, 0.5)) == rec["(100,0.5)"]
            continue
        plan = otome.bipartite_soft_matching_plan(rec["k"], rec["r"])
        assert torch.equal(otome.merge_mean(plan, rec["k"]), rec["k_merged"]), name
        assert torch.equal(otome.merge_mean(plan, rec["v"]), rec["v_merged"]), name
    # edge cases the reference handles: r == 0 / r larger than half (clamped) / odd token counts
    k = torch.randn(1, 9, 8, generator=torch.Generator().manual_seed(0))
    assert otome.bipartite_soft_matching_plan(k, 0) is None
    plan = otome.bipartite_soft_matching_plan(k, 100)
    assert otome.merge_mean(plan, k).shape == (1, 9 - 4, 8)
    assert otome.parse_r(4, [3]) == [3, 0, 0, 0]
    assert otome.parse_r(16, 8) == [8] * 16


def test_oracle_unet_vae_pipeline_fixtures():
    g = torch.load(os.path.join(GOLD, "oracle_tiny.pt"))
    cfg = UNetConfig.tiny()
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    u = g["unet_tiny"]
    with torch.no_grad():
        assert (unet_forward(P, cfg, u["x"], u["t"], u["ctx"]) - u["eps"]).abs().max().item() < 1e-4
        t = g["unet_tiny_tome"]
        assert (unet_forward(P, cfg, u["x"], u["t"], u["ctx"], tome_r=t["r"]) - t["eps"]).abs().max().item() < 1e-4
        vcfg = VAEConfig.tiny()
        VP = synth_params(vae_param_shapes(vcfg), seed=4321)
        assert (vae_decode(VP, vcfg, g["vae_tiny"]["z"]) - g["vae_tiny"]["img"]).abs().max().item() < 1e-4
        e = g["vae_tiny_enc"]
        assert (vae_encode_moments(VP, vcfg, e["img"]) - e["moments"]).abs().max().item() < 1e-4
        unet = OracleUNet(cfg, P)
        emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=torch.Generator().manual_seed(11))
        unc = torch.randn(1, 77, cfg.cross_attention_dim, generator=torch.Generator().manual_seed(12)).expand(2, -1, -1)
        cfgu = osamp.CFGParallel(unet, unc, emb, 7.5)
        rec = g["pipe_tiny/euler_a"]
        lat = osamp.txt2img_latents(cfgu, batch=2, in_channels=4, height=128, width=128, sample_size=16,
                                    seeds=[420420420, 420420421], steps=rec["steps"], sampler="euler_a")
        assert (lat - rec["latents"]).abs().max().item() < 1e-3 * rec["latents"].abs().max().item()


def test_c1_fixture_present_and_sane():
    """The full-size C1 latents (SD1.5 arch, 10 DDIM steps, seed 420420420) are what the GPU parity test of
    BASELINE config 1 compares against."""
    rec = torch.load(os.path.join(GOLD, "c1_sd15_ddim10.pt"))
    assert rec["latents"].shape == (1, 4, 64, 64) and rec["steps"] == 10 and rec["sampler"] == "ddim"
    assert torch.isfinite(rec["latents"]).all()


def test_oracle_controlnet_residual_semantics():
    """`down_block_additional_residuals` / `mid_block_additional_residual` as gyre/pipeline/unet/core.py:213-239 hands
    them to the UNet: zero residuals change nothing, the down path does not see them, the mid output moves by exactly
    the mid residual, and every skip residual reaches the output."""
    cfg = UNetConfig.tiny()
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(1, 4, 16, 16, generator=gen)
    ctx = torch.randn(1, 77, cfg.cross_attention_dim, generator=gen)
    taps0, taps1 = {}, {}
    with torch.no_grad():
        base = unet_forward(P, cfg, x, 300, ctx, taps=taps0)
        shapes = [taps0["conv_in"].shape]
        h = w = 16
        for i, c in enumerate(cfg.block_out_channels):
            shapes += [(1, c, h, w)] * cfg.layers_per_block
            if i < len(cfg.block_out_channels) - 1:
                h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
                shapes.append((1, c, h, w))
        zeros = [torch.zeros(s) for s in shapes]
        mid0 = torch.zeros_like(taps0["mid_block"])
        same = unet_forward(P, cfg, x, 300, ctx, down_block_additional_residuals=zeros, mid_block_additional_residual=mid0)
        assert torch.equal(same, base)
        down = [torch.randn(s, generator=gen) * 0.3 for s in shapes]
        mid = torch.randn(mid0.shape, generator=gen) * 0.3
        out = unet_forward(P, cfg, x, 300, ctx, taps=taps1, down_block_additional_residuals=down,
                           mid_block_additional_residual=mid)
    for k in taps0:
        if k.startswith("down_blocks") or k == "conv_in":
            assert torch.equal(taps0[k], taps1[k]), f"{k}: the down path must not see the residuals"
    assert torch.allclose(taps1["mid_block"] - taps0["mid_block"], mid, atol=1e-5)
    assert (out - base).abs().max() > 1e-3
    # each skip residual individually reaches the output
    for k in range(len(shapes)):
        one = [d if j == k else z for j, (d, z) in enumerate(zip(down, zeros))]
        with torch.no_grad():
            o = unet_forward(P, cfg, x, 300, ctx, down_block_additional_residuals=one)
        assert (o - base).abs().max() > 1e-5, f"skip residual {k} has no effect"
    with pytest.raises(AssertionError):
        unet_forward(P, cfg, x, 300, ctx, down_block_additional_residuals=zeros[:-1])


def test_oracle_t2i_adapter_semantics():
    """adapter_states: in-place add before each block's downsampler (unet_patcher.py:21-60): the block's last skip and
    the downsampler input both move by the state; earlier layers of the block do not."""
    cfg = UNetConfig.tiny()
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(1, 4, 16, 16, generator=gen)
    ctx = torch.randn(1, 77, cfg.cross_attention_dim, generator=gen)
    states, h = [], 16
    for c in cfg.block_out_channels:
        states.append(torch.randn(1, c, h, h, generator=gen) * 0.3)
        h = (h - 1) // 2 + 1
    t0, t1 = {}, {}
    with torch.no_grad():
        base = unet_forward(P, cfg, x, 300, ctx, taps=t0)
        assert torch.equal(unet_forward(P, cfg, x, 300, ctx, adapter_states=[torch.zeros_like(s) for s in states]), base)
        assert torch.equal(unet_forward(P, cfg, x, 300, ctx, adapter_states=[]), base)
        out = unet_forward(P, cfg, x, 300, ctx, taps=t1, adapter_states=states)
    # level 0 layers run before the first state is added: untouched; level 1 sees the level-0 state through the downsampler
    for j in range(cfg.layers_per_block):
        assert torch.equal(t0[f"down_blocks.0.resnets.{j}"], t1[f"down_blocks.0.resnets.{j}"])
    assert not torch.equal(t0["down_blocks.1.resnets.0"], t1["down_blocks.1.resnets.0"])
    assert (out - base).abs().max() > 1e-3
    with pytest.raises(AssertionError):
        unet_forward(P, cfg, x, 300, ctx, adapter_states=states[:-1])


# ---- RNG contract + CFG / embedding / extra-channel wrappers vs the reference's own classes (wrappers.pt)
def _wrappers():
    return torch.load(os.path.join(GOLD, "wrappers.pt"))


@pytest.mark.parametrize("dt_name,dt", [("fp32", torch.float32), ("fp16", torch.float16)])
def test_batched_randn_matches_reference_randtools(dt_name, dt):
    """gyre/pipeline/randtools.py:39-64 imported by scripts/make_golden.py: oracle AND product host code, bit-exact."""
    from gyre_b200 import randtools as ours
    W = _wrappers()
    seeds = [420420420, 420420421, 7]
    go = [torch.Generator("cpu").manual_seed(s) for s in seeds]
    gb = [torch.Generator("cpu").manual_seed(s) for s in seeds]
    for call in range(3):
        ref = W[f"randn_{dt_name}_{call}"]
        assert torch.equal(osamp.batched_randn((3, 4, 8, 8), go, "cpu", dt), ref)
        assert torch.equal(ours.batched_randn((3, 4, 8, 8), gb, torch.device("cpu"), dt), ref)
    g2o = [torch.Generator("cpu").manual_seed(5), torch.Generator("cpu").manual_seed(6)]
    g2b = [torch.Generator("cpu").manual_seed(5), torch.Generator("cpu").manual_seed(6)]
    assert torch.equal(osamp.batched_randn((4, 4, 4, 4), g2o, "cpu", dt), W[f"randn_cycled_{dt_name}"])
    assert torch.equal(ours.batched_randn((4, 4, 4, 4), g2b, torch.device("cpu"), dt), W[f"randn_cycled_{dt_name}"])
    with pytest.raises(ValueError):
        ours.batched_randn((3, 4, 4, 4), g2b, torch.device("cpu"), dt)


@pytest.mark.parametrize("dt_name,dt", [("fp32", torch.float32), ("fp16", torch.float16)])
@pytest.mark.parametrize("t_name", ["tvec", "tint"])
def test_cfg_wrapper_stack_matches_reference(dt_name, dt, t_name):
    """CFGUNet_Parallel / CFGUNet_Sequential over UNetWithEmbeddings over CFGUNetFromDiffusersUNet
    (gyre/pipeline/unet/cfg.py:26-57, core.py:242-274), as composed at unified_pipeline.py:2238,2326-2337."""
    from fakes import FakeDiffusersUNet
    W = _wrappers()
    I = W["inputs"]
    t = I["t_vec"] if t_name == "tvec" else I["t_int"]
    ocfg = osamp.CFGParallel(FakeDiffusersUNet(), I["unc"].to(dt), I["cond"].to(dt), I["scale"])
    got = ocfg(I["lat"].to(dt), t)
    assert torch.equal(got, W[f"cfg_{dt_name}_plain_{t_name}_parallel"])
    # the sequential execution mode is the same function up to the rounding of two separate UNet calls
    seq = W[f"cfg_{dt_name}_plain_{t_name}_sequential"]
    assert torch.allclose(got.float(), seq.float(), atol=2e-2 if dt == torch.float16 else 1e-5)


def test_controlnet_oracle_matches_reference_fixture():
    """tests/golden/controlnet.pt: per-tensor sums of what the reference's in-tree ControlNetModel.forward returned (diffusers
    blocks stood in for by the oracle's; scripts/make_golden.py:pin_controlnet asserted bit-equality tensor by tensor)."""
    import os
    import torch
    from oracle import controlnet as ocn
    from oracle.unet import UNetConfig, synth_params
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "controlnet.pt"))
    cfg = UNetConfig.tiny()
    for order in ("rgb", "bgr"):
        g = G[order]
        P = synth_params(ocn.controlnet_param_shapes(cfg), seed=g["weights_seed"])
        cond = g["cond"].float()
        cond = torch.flip(cond, dims=[1]) if order == "bgr" else cond
        for tn, t in g["t"].items():
            with torch.no_grad():
                down, mid = ocn.controlnet_forward(P, cfg, g["x"], t, g["ctx"], cond)
            sums = [float(d.double().sum()) for d in down] + [float(mid.double().sum())]
            for a, b in zip(sums, g["sums"][tn]):
                assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (order, tn)


def test_attention_module_oracle_matches_reference_fixture():
    """tests/golden/attention.pt: outputs of the REFERENCE's MemoryEfficientCrossAttention / ToMeMemoryEfficientCrossAttention
    modules (scripts/make_golden.py:pin_attention; real vendored tome.merge inside the ToMe variant)."""
    import os
    import torch
    from oracle.unet import attention
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "attention.pt"))
    assert set(G) == {"self", "cross", "self_d40", "tome_r16", "tome_half", "tome_odd"}
    for name, g in G.items():
        C_, heads, N_, ctx_dim, L, r = g["config"]
        P = {f"a.{k}": v for k, v in g["state_dict"].items()}
        with torch.no_grad():
            out = attention(P, "a", g["x"], g["ctx"], heads, tome_r=r)
        assert (out - g["out"]).abs().max().item() <= 2e-6, name
