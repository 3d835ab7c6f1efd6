"""Host-side logic of the product package, on the CPU: schedule / scalar maths, RNG contract, weight
inventory, sharding helpers, and the loud failure when no GPU is present (there is NO CPU fallback)."""
import pytest
import torch

from gyre_b200 import common_scheduler as cs
from gyre_b200 import dist as gdist
from gyre_b200.config import UNetConfig, VAEConfig
from gyre_b200.pipeline import generate_latents
from gyre_b200.randtools import batched_randn, predraw_noise
from gyre_b200.unet import parse_r
from gyre_b200.weights import synth_state_dict, unet_param_shapes, vae_param_shapes
from oracle import sampling as osamp
from oracle import tome as otome
from oracle import unet as ounet
from oracle import vae as ovae


def test_schedule_matches_oracle():
    sch = cs.DiscreteSchedule(cs.sd_alphas_cumprod())
    den = osamp.EpsDenoiser(lambda x, t: x, osamp.sd_alphas_cumprod())
    assert torch.equal(sch.sigmas, den.sigmas)
    for n in (7, 10, 50):
        t = torch.linspace(999, 0, n)
        mine = torch.cat([sch.t_to_sigma(t), torch.zeros(1)])
        assert torch.equal(mine, osamp.k_sigmas(den, n))
    s = osamp.k_sigmas(den, 50)[:-1].half().float()
    assert torch.equal(sch.sigma_to_t(s), den.sigma_to_t(s))
    assert torch.equal(cs.get_sigmas_karras(11, sch.sigma_min, sch.sigma_max, 7.0),
                       osamp.get_sigmas_karras(11, den.sigma_min, den.sigma_max, 7.0))
    a = cs.get_ancestral_step(torch.tensor(3.0), torch.tensor(2.0))
    b = osamp.get_ancestral_step(torch.tensor(3.0), torch.tensor(2.0))
    assert float(a[0]) == float(b[0]) and float(a[1]) == float(b[1])
    assert cs.get_ancestral_step(torch.tensor(3.0), torch.tensor(2.0), eta=0.0)[1] == 0.0


def _dummy_guided():
    from gyre_b200.cfg import B200GuidedUNet
    g = object.__new__(B200GuidedUNet)
    g.guidance_scale = 7.5
    return g


def test_kdiffusion_scheduler_host_state():
    sched = cs.build_scheduler("k_euler_ancestral", [torch.Generator().manual_seed(1)], "cpu", torch.float16)
    with pytest.raises(ValueError):
        sched.set_timesteps(10)                      # eps unet must be set first (common_scheduler.py:437-438)
    with pytest.raises(TypeError):
        sched.set_eps_unets([lambda x, t: x])
    sched.set_eps_unets([_dummy_guided()])
    sched.set_timesteps(50)
    den = osamp.EpsDenoiser(lambda x, t: x, osamp.sd_alphas_cumprod())
    assert torch.equal(sched.sigmas, osamp.k_sigmas(den, 50))
    x = torch.ones(1, 4, 8, 8)
    assert torch.equal(sched.prepare_initial_latents(x), x * sched.sigmas[0])
    with pytest.raises(RuntimeError):
        sched.set_eps_unets([_dummy_guided()])       # "Can't set eps_unet once set_timesteps has been called"
    sched2 = cs.build_scheduler("k_euler_ancestral", [torch.Generator()], "cpu", torch.float16)
    sched2.set_eps_unets([_dummy_guided()])
    sched2.set_timesteps(20, strength=0.5)
    assert sched2.start_offset == 10
    with pytest.raises(ValueError):
        sched2.set_timesteps(20, strength=0.5, start_offset=3)
    sched3 = cs.build_scheduler("k_euler", [torch.Generator()], "cpu", torch.float16)
    sched3.set_eps_unets([_dummy_guided()])
    sched3.set_timesteps(8, config=cs.SchedulerConfig(karras_rho=7.0))
    assert torch.equal(sched3.sigmas, osamp.get_sigmas_karras(8, den.sigma_min, den.sigma_max, 7.0))
    with pytest.raises(NotImplementedError):
        cs.build_scheduler("dpm_solver_pp_2", [torch.Generator()], "cpu", torch.float16)
    # the loop needs CUDA tensors: no silent CPU path
    with pytest.raises(Exception):
        sched.loop(torch.zeros(1, 4, 8, 8))


class _CpuEngine:
    """Stand-in for KDiffusionScheduler._Engine: the two device kernels replaced by their definitions
    (den = x + eps * c_out ; out = sum coef * tensor), so the HOST side of the generic sampler loops - every
    coefficient, branch and noise draw - can be checked on the CPU against the vendored k-diffusion results."""

    def __init__(self, sched, latents, eps_fn):
        self.s, self.shape = sched, tuple(latents.shape)
        self.den = osamp.EpsDenoiser(eps_fn, osamp.sd_alphas_cumprod())

    def denoise(self, x, sigma):
        return self.den(x, torch.as_tensor(sigma, dtype=torch.float32) * x.new_ones([x.shape[0]]))

    def lin(self, terms, out=None):
        acc = None
        for c, t in terms:
            if t is None:
                continue
            term = float(c) * t
            acc = term if acc is None else acc + term
        return acc

    def noise(self):
        return batched_randn(self.shape, self.s.generators, "cpu", self.s.dtype).float()

    def err_norm(self, x_low, x_high, x_prev, atol, rtol):
        delta = torch.maximum(torch.tensor(atol), torch.tensor(rtol) * torch.maximum(x_low.abs(), x_prev.abs()))
        return float(torch.linalg.norm((x_low - x_high) / delta) / x_low.numel() ** 0.5)


def _toy_eps(x, t):
    tt = t.float().reshape(-1, *([1] * (x.ndim - 1)))
    return 0.7 * torch.tanh(x) + 0.001 * tt * x.roll(1, -1)


@pytest.mark.parametrize("enum_name,gold_name", [("k_heun", "heun"), ("k_dpm_2", "dpm_2"), ("k_dpm_2_ancestral", "dpm_2_a"),
                                                 ("k_lms", "lms"), ("k_dpmpp_2s_ancestral", "dpmpp_2s_a"),
                                                 ("k_dpmpp_sde", "dpmpp_sde"), ("k_dpmpp_2m", "dpmpp_2m")])
@pytest.mark.parametrize("steps", [7, 20])
@pytest.mark.parametrize("dtype_name", ["fp32", "fp16"])
def test_generic_sampler_host_logic_vs_vendored_golden(enum_name, gold_name, steps, dtype_name):
    import os
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "samplers.pt"))[f"{gold_name}/{steps}/{dtype_name}"]
    ldt = torch.float32 if dtype_name == "fp32" else torch.float16
    gens = [torch.Generator("cpu").manual_seed(sd) for sd in gold["seeds"]]
    sched = cs.build_scheduler(enum_name, gens, "cpu", ldt)
    sched.set_eps_unets([_dummy_guided()])
    sched.set_timesteps(steps)
    assert torch.equal(sched.sigmas, gold["sigmas"])
    x0 = sched.prepare_initial_latents(batched_randn(gold["shape"], gens, "cpu", ldt)).float()
    sched._make_engine = lambda latents: _CpuEngine(sched, latents, _toy_eps)
    sigmas = sched.sigmas.to(ldt).float()
    out = sched._loop_generic(x0, sigmas, lambda it: it, torch.float32, 1.0)
    err = (out - gold["result"]).abs().max().item()
    scale = gold["result"].abs().max().item()
    # same linear maps as the vendored loops, coefficients folded differently: fp32 rounding noise only
    assert err <= 2e-5 * max(scale, 1.0), f"{enum_name}/{steps}/{dtype_name}: {err} (scale {scale})"


@pytest.mark.parametrize("enum_name,gold_name", [("k_euler", "euler"), ("k_heun", "heun"), ("k_dpm_2", "dpm_2")])
def test_churn_host_logic_vs_vendored_golden(enum_name, gold_name):
    """s_churn > 0 (sampling.py:124-129): gamma / sigma_hat / the extra noise injection, against the vendored loops."""
    import os
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "samplers.pt"))[f"{gold_name}_churn/12/fp32"]
    gens = [torch.Generator("cpu").manual_seed(sd) for sd in gold["seeds"]]
    sched = cs.build_scheduler(enum_name, gens, "cpu", torch.float32)
    sched.set_eps_unets([_dummy_guided()])
    churn, tmin, tmax = gold["churn"]
    sched.set_timesteps(12, config=cs.SchedulerConfig(churn=churn, churn_tmin=tmin, churn_tmax=tmax))
    x0 = sched.prepare_initial_latents(batched_randn(gold["shape"], gens, "cpu", torch.float32)).float()
    sched._make_engine = lambda latents: _CpuEngine(sched, latents, _toy_eps)
    out = sched._loop_generic(x0, sched.sigmas.float(), lambda it: it, torch.float32, 1.0)
    err = (out - gold["result"]).abs().max().item()
    assert err <= 2e-5 * max(gold["result"].abs().max().item(), 1.0), f"{enum_name} churn: {err}"


@pytest.mark.parametrize("key,eta", [("dpm_fast/7/fp32", None), ("dpm_fast/12/fp32", None), ("dpm_fast/20/fp32", None),
                                     ("dpm_fast/7/fp16", None), ("dpm_fast/12/fp16", None), ("dpm_fast/20/fp16", None),
                                     ("dpm_fast/11/fp32/eta0.6", 0.6), ("dpm_fast/11/fp16/eta0.6", 0.6)])
def test_dpm_fast_host_logic_vs_vendored_golden(key, eta):
    """`sample_dpm_fast` (sampling.py:482-491): step orders (3...3,2,1 | 3...3,n%3), the t = -log sigma grid built from
    the dtype-cast sigma_min / sigma_max, the three solver stages as linear maps, eta > 0 and the per-step noise draw."""
    import os
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "samplers.pt"))[key]
    ldt = torch.float16 if "/fp16" in key else torch.float32
    gens = [torch.Generator("cpu").manual_seed(sd) for sd in gold["seeds"]]
    sched = cs.build_scheduler("dpm_fast", gens, "cpu", ldt)
    sched.set_eps_unets([_dummy_guided()])
    sched.set_timesteps(gold["steps"], config=cs.SchedulerConfig(eta=eta))
    assert torch.equal(sched.sigmas, gold["sigmas"])
    x0 = sched.prepare_initial_latents(batched_randn(gold["shape"], gens, "cpu", ldt)).float()
    sched._make_engine = lambda latents: _CpuEngine(sched, latents, _toy_eps)
    sigmas = sched.sigmas.to(ldt).float()
    out = sched._loop_dpm_fast(x0, sigmas, lambda it: it, torch.float32, 0.0 if eta is None else eta)
    err = (out - gold["result"]).abs().max().item()
    scale = gold["result"].abs().max().item()
    print(f"{key}: host-logic max abs err {err:.3e} (scale {scale:.2f})")
    assert err <= 5e-5 * max(scale, 1.0), f"{key}: {err}"


@pytest.mark.parametrize("key,eta", [("dpm_adaptive/10/fp32", None), ("dpm_adaptive/25/fp32", None),
                                     ("dpm_adaptive/10/fp16", None), ("dpm_adaptive/25/fp16", None),
                                     ("dpm_adaptive/10/fp32/eta0.5", 0.5), ("dpm_adaptive/10/fp16/eta0.5", 0.5)])
def test_dpm_adaptive_host_logic_vs_vendored_golden(key, eta):
    """`sample_dpm_adaptive` (sampling.py:494-506): the embedded 2/3 pair, the PID controller, accept / reject, the time
    variable kept in the latent dtype, noise on accepted steps only - same number of trial steps as the vendored run."""
    import os
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "samplers.pt"))[key]
    ldt = torch.float16 if "/fp16" in key else torch.float32
    gens = [torch.Generator("cpu").manual_seed(sd) for sd in gold["seeds"]]
    sched = cs.build_scheduler("dpm_adaptive", gens, "cpu", ldt)
    sched.set_eps_unets([_dummy_guided()])
    sched.set_timesteps(gold["steps"], config=cs.SchedulerConfig(eta=eta))
    x0 = sched.prepare_initial_latents(batched_randn(gold["shape"], gens, "cpu", ldt)).float()
    sched._make_engine = lambda latents: _CpuEngine(sched, latents, _toy_eps)
    sigmas = sched.sigmas.to(ldt).float()
    out = sched._loop_dpm_adaptive(x0, sigmas, lambda it: it, torch.float32, 0.0 if eta is None else eta)
    assert sched.last_solver_info["steps"] == gold["info"]["steps"]
    err = (out - gold["result"]).abs().max().item()
    scale = gold["result"].abs().max().item()
    print(f"{key}: host-logic max abs err {err:.3e} (scale {scale:.2f}), {gold['info']}")
    assert err <= 5e-5 * max(scale, 1.0), f"{key}: {err}"


def test_ddim_scheduler_host_state():
    sched = cs.build_scheduler("ddim", [torch.Generator()], "cpu", torch.float32)
    sched.set_eps_unets([_dummy_guided()])
    sched.set_timesteps(10)
    assert sched.timesteps.tolist() == osamp.ddim_timesteps(10).tolist()
    assert sched.init_noise_sigma == 1.0
    x = torch.randn(1, 4, 8, 8)
    assert torch.equal(sched.prepare_initial_latents(x), x)
    assert torch.equal(sched.scale_latents(x, 5), x)


def test_rng_contract():
    seeds = [420420420, 7, 99]
    g1 = [torch.Generator().manual_seed(s) for s in seeds]
    g2 = [torch.Generator().manual_seed(s) for s in seeds]
    a = batched_randn([3, 4, 8, 8], g1, "cpu", torch.float16)
    b = osamp.batched_randn([3, 4, 8, 8], g2, "cpu", torch.float16)
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        batched_randn([4, 4, 8, 8], g1, "cpu", torch.float16)
    # drawing a run's noise up front == per-step draws (generators advance identically)
    g1 = [torch.Generator().manual_seed(s) for s in seeds]
    g2 = [torch.Generator().manual_seed(s) for s in seeds]
    pre = predraw_noise(5, (3, 4, 8, 8), g1, "cpu", torch.float16)
    for i in range(5):
        assert torch.equal(pre[i], batched_randn([3, 4, 8, 8], g2, "cpu", torch.float16))
    assert torch.equal(torch.randn(3, generator=g1[0]), torch.randn(3, generator=g2[0]))
    # batch independence of the draws (reference tests/batch_independance.py): sample i only depends on seed i
    g3 = [torch.Generator().manual_seed(seeds[1])]
    solo = batched_randn([1, 4, 8, 8], g3, "cpu", torch.float16)
    g4 = [torch.Generator().manual_seed(s) for s in seeds]
    assert torch.equal(batched_randn([3, 4, 8, 8], g4, "cpu", torch.float16)[1:2], solo)
    assert predraw_noise(0, (3, 4, 8, 8), g1, "cpu", torch.float16).shape[0] == 0


@pytest.mark.parametrize("hw,ss", [((512, 512), 64), ((256, 384), 64), ((768, 600), 64), ((512, 512), 96)])
def test_initial_latents_follow_generateLatents(hw, ss):
    """Txt2imgMode.generateLatents crop / insert rule, checked against the oracle's restatement."""
    h, w = hw
    seeds = [1, 2]
    g1 = [torch.Generator().manual_seed(s) for s in seeds]
    lat = generate_latents(g1, 2, 4, h, w, ss, "cpu", torch.float32)
    assert lat.shape == (2, 4, h // 8, w // 8)
    # oracle path: run txt2img_latents with a 0-step "sampler" is not available, so restate through its helper
    g2 = [torch.Generator().manual_seed(s) for s in seeds]
    mid = osamp.batched_randn([2, 4, ss, ss], g2, "cpu", torch.float32)
    hh, ww = h // 8, w // 8
    o2, o3 = (ss - hh) // 2, (ss - ww) // 2
    if o2 > 0:
        mid = mid[:, :, o2:o2 + hh, :]
    if o3 > 0:
        mid = mid[:, :, :, o3:o3 + ww]
    if o2 >= 0 and o3 >= 0:
        ref = mid
    else:
        ref = osamp.batched_randn((2, 4, hh, ww), g2, "cpu", torch.float32)
        p2, p3 = (ref.shape[2] - mid.shape[2]) // 2, (ref.shape[3] - mid.shape[3]) // 2
        ref[:, :, p2:p2 + mid.shape[2], p3:p3 + mid.shape[3]] = mid
    assert torch.equal(lat, ref)


def test_parse_r_matches_reference_semantics():
    for arg in (8, (8, -1.0), (100, 0.5), [3, 2, 1]):
        a = parse_r(16, list(arg) if isinstance(arg, list) else arg)
        b = otome.parse_r(16, list(arg) if isinstance(arg, list) else arg)
        assert a == b
    assert parse_r(16, int(0.5)) == [0] * 16      # the reference's literal `int(value)` (SURVEY finding 3)


def test_weight_inventory_matches_oracle():
    for mine, theirs in ((UNetConfig.tiny(), ounet.UNetConfig.tiny()), (UNetConfig.sd15(), ounet.UNetConfig.sd15()),
                         (UNetConfig.sd21_v(), ounet.UNetConfig.sd21_v()),
                         (UNetConfig.sd15_inpaint(), ounet.UNetConfig.sd15_inpaint())):
        assert unet_param_shapes(mine) == ounet.unet_param_shapes(theirs)
    assert vae_param_shapes(VAEConfig.sd()) == ovae.vae_param_shapes(ovae.VAEConfig.sd())
    assert vae_param_shapes(VAEConfig.tiny()) == ovae.vae_param_shapes(ovae.VAEConfig.tiny())
    sh = unet_param_shapes(UNetConfig.tiny())
    a, b = synth_state_dict(sh, 1234), ounet.synth_params(sh, 1234)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    n_params = sum(torch.Size(s).numel() for s in unet_param_shapes(UNetConfig.sd15()).values())
    assert 855e6 < n_params < 865e6                 # "~860 M params" 
    assert UNetConfig.from_any(ounet.UNetConfig.sd21_v()).num_heads == (5, 10, 20, 20)


def test_shard_range():
    for total in (1, 7, 8, 16, 64):
        for ws in (1, 2, 4, 8):
            spans = [gdist.shard_range(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gyre_b200 import _native as N
    from gyre_b200.unet import B200UNet
    from gyre_b200.vae import B200VAE
    with pytest.raises(N.NativeError):
        B200UNet(UNetConfig.tiny())
    with pytest.raises(N.NativeError):
        B200VAE(VAEConfig.tiny())
    with pytest.raises(N.NativeError):
        N.gemm(torch.zeros(8, 8).half(), torch.zeros(8, 8).half())


def _diffusers_unet_config(**kw):
    """What `UNet2DConditionModel.config` looks like in diffusers 0.16 (the names gyre's ckpt_utils.py:280-304 writes)."""
    d = dict(act_fn="silu", attention_head_dim=8, block_out_channels=[320, 640, 1280, 1280], center_input_sample=False,
             cross_attention_dim=768, down_block_types=["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"],
             up_block_types=["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3, downsample_padding=1, flip_sin_to_cos=True,
             freq_shift=0, in_channels=4, layers_per_block=2, mid_block_scale_factor=1, norm_eps=1e-5,
             norm_num_groups=32, out_channels=4, sample_size=64, only_cross_attention=False, dual_cross_attention=False,
             class_embed_type=None, num_class_embeds=None, use_linear_projection=False, upcast_attention=False)
    d.update(kw)
    return d


def test_unet_config_from_diffusers_names():
    """a diffusers config must map attention_head_dim -> heads and down_block_types -> attention levels
    (an SD2.x config used to get 8 heads per level silently)."""
    sd15 = UNetConfig.from_any(_diffusers_unet_config())
    assert sd15.num_heads == (8, 8, 8, 8) and sd15.attn_levels == (True, True, True, False)
    sd21 = UNetConfig.from_any(_diffusers_unet_config(attention_head_dim=[5, 10, 20, 20], cross_attention_dim=1024,
                                                      use_linear_projection=True, upcast_attention=True,
                                                      sample_size=96, prediction_type="v_prediction"))
    assert sd21.num_heads == (5, 10, 20, 20) and sd21.use_linear_projection and sd21.upcast_attention
    assert sd21 == UNetConfig.sd21_v()
    sdxl = UNetConfig.from_any(dict(block_out_channels=[320, 640, 1280], attention_head_dim=[5, 10, 20],
                                    down_block_types=["DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"],
                                    up_block_types=["CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"],
                                    transformer_layers_per_block=[1, 2, 10], cross_attention_dim=2048,
                                    use_linear_projection=True, sample_size=128, addition_embed_type="text_time",
                                    addition_time_embed_dim=256, projection_class_embeddings_input_dim=2816))
    assert sdxl == UNetConfig.sdxl()

    class Obj:      # attribute-style config (diffusers FrozenDict also allows attribute access)
        pass
    o = Obj()
    for k, v in _diffusers_unet_config(attention_head_dim=[5, 10, 20, 20]).items():
        setattr(o, k, v)
    assert UNetConfig.from_any(o).num_heads == (5, 10, 20, 20)
    # the oracle's config (our own field names) still round-trips
    assert UNetConfig.from_any(ounet.UNetConfig.tiny()).num_heads == UNetConfig.tiny().num_heads
    with pytest.raises(ValueError):          # heads unresolvable
        UNetConfig.from_any(dict(block_out_channels=[320, 640, 1280, 1280],
                                 down_block_types=["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"]))
    with pytest.raises(ValueError):          # attention levels unresolvable
        UNetConfig.from_any(dict(block_out_channels=[320, 640, 1280, 1280], attention_head_dim=8))
    with pytest.raises(ValueError):          # 320 channels do not split into 7 heads
        UNetConfig.from_any(_diffusers_unet_config(attention_head_dim=7))
    with pytest.raises(NotImplementedError):
        UNetConfig.from_any(_diffusers_unet_config(only_cross_attention=True))
    with pytest.raises(NotImplementedError):
        UNetConfig.from_any(_diffusers_unet_config(down_block_types=["AttnDownBlock2D"] * 4))
