#!/usr/bin/env python
"""Generate tests/golden/*.pt (run in the build container, where /root/reference is mounted).

Part 1 PINS the oracle: it imports the reference's own vendored k-diffusion, ToMe and DDIM sources,
runs them on seeded inputs, asserts that the oracle restatement reproduces them, and stores the
REFERENCE outputs as golden vectors.
Part 2 stores oracle outputs for the UNet / VAE / pipeline (no reference implementation of those is
importable: diffusers is an absent third-party dependency -> "parity unpinned" there).

    python scripts/make_golden.py [--full]     (--full adds the SD1.5-size C1 fixture, ~5 min of CPU)
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _vendored  # noqa: E402
from oracle import sampling as osamp  # noqa: E402
from oracle import tome as otome  # noqa: E402
from oracle.unet import UNetConfig, OracleUNet, synth_params, unet_forward, unet_param_shapes  # noqa: E402
from oracle.vae import VAEConfig, OracleVAE, vae_decode, vae_encode_moments, vae_param_shapes  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def toy_eps(x, t):
    t = t if torch.is_tensor(t) else torch.tensor(t)
    tt = t.float().reshape(-1, *([1] * (x.ndim - 1)))
    return 0.7 * torch.tanh(x) + 0.001 * tt * x.roll(1, -1)


def pin_samplers():
    _, ks, kext = _vendored.k_diffusion()
    dpmpp = _vendored.gyre_dpmpp_2m()
    acp = osamp.sd_alphas_cumprod()
    out = {}
    for dtype_name, ldt in (("fp32", torch.float32), ("fp16", torch.float16)):
        for name in ("euler_a", "euler", "heun", "dpmpp_2m", "dpm_2", "dpm_2_a", "lms", "dpmpp_2s_a", "dpmpp_sde"):
            for steps in (7, 20):
                shape = (2, 4, 8, 8)
                seeds = [420420420, 420420421]
                # ---- reference (vendored) run, following KDiffusionScheduler.set_timesteps/loop
                gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]
                ref_den = kext.DiscreteEpsDDPMDenoiser(toy_eps, acp, quantize=True)
                t = torch.linspace(len(ref_den.sigmas) - 1, 0, steps)
                sig_full = ks.append_zero(ref_den.t_to_sigma(t))
                x0 = osamp.batched_randn(shape, gens, "cpu", ldt) * sig_full[0]
                sig = sig_full.to(ldt).float()
                x0 = x0.float()
                ns = lambda *_: osamp.batched_randn(shape, gens, "cpu", ldt).float()
                orig_randn_like = ks.torch.randn_like
                if name == "euler_a":
                    ref = ks.sample_euler_ancestral(ref_den, x0, sig, noise_sampler=ns, disable=True)
                elif name == "dpm_2_a":
                    ref = ks.sample_dpm_2_ancestral(ref_den, x0, sig, noise_sampler=ns, disable=True)
                elif name == "dpmpp_2s_a":
                    ref = ks.sample_dpmpp_2s_ancestral(ref_den, x0, sig, noise_sampler=ns, disable=True)
                elif name == "dpmpp_sde":
                    ref = ks.sample_dpmpp_sde(ref_den, x0, sig, noise_sampler=ns, disable=True)
                elif name == "lms":
                    ref = ks.sample_lms(ref_den, x0, sig, disable=True)
                elif name in ("euler", "heun", "dpm_2"):
                    class _T:  # TorchRandOverride (randtools.py:67-90) restated for the patch
                        def __getattr__(self, k):
                            return getattr(torch, k)

                        def randn_like(self, inp, **kw):
                            return ns()
                    ks.torch = _T()
                    try:
                        fn = {"euler": ks.sample_euler, "heun": ks.sample_heun, "dpm_2": ks.sample_dpm_2}[name]
                        ref = fn(ref_den, x0, sig, disable=True)
                    finally:
                        ks.torch = torch
                else:
                    ref = dpmpp.sample_dpmpp_2m(ref_den, x0, sig, disable=True, warmup_lms=True, ddim_cutoff=0.1)
                assert ks.torch.randn_like is orig_randn_like
                # ---- oracle restatement
                gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]
                den = osamp.EpsDenoiser(toy_eps, acp)
                sig2 = osamp.k_sigmas(den, steps)
                assert torch.equal(sig2, sig_full), "sigma schedule mismatch"
                y0 = (osamp.batched_randn(shape, gens, "cpu", ldt) * sig2[0]).float()
                ns2 = lambda *_: osamp.batched_randn(shape, gens, "cpu", ldt).float()
                s2 = sig2.to(ldt).float()
                if name == "euler_a":
                    got = osamp.sample_euler_ancestral(den, y0, s2, ns2)
                elif name == "euler":
                    got = osamp.sample_euler(den, y0, s2, lambda x: ns2())
                elif name == "heun":
                    got = osamp.sample_heun(den, y0, s2, lambda x: ns2())
                elif name == "dpm_2":
                    got = osamp.sample_dpm_2(den, y0, s2, lambda x: ns2())
                elif name == "dpm_2_a":
                    got = osamp.sample_dpm_2_ancestral(den, y0, s2, ns2)
                elif name == "lms":
                    got = osamp.sample_lms(den, y0, s2)
                elif name == "dpmpp_2s_a":
                    got = osamp.sample_dpmpp_2s_ancestral(den, y0, s2, ns2)
                elif name == "dpmpp_sde":
                    got = osamp.sample_dpmpp_sde(den, y0, s2, ns2)
                else:
                    got = osamp.sample_dpmpp_2m(den, y0, s2, warmup_lms=True, ddim_cutoff=0.1)
                err = (got - ref).abs().max().item()
                assert err == 0.0, f"{name}/{steps}/{dtype_name}: oracle != vendored k-diffusion ({err})"
                out[f"{name}/{steps}/{dtype_name}"] = {"seeds": seeds, "shape": shape, "sigmas": sig_full,
                                                        "result": ref}
    # DPM-Solver-Fast: takes (sigma_min, sigma_max, n) instead of a sigma list (common_scheduler.py:590-594)
    for dtype_name, ldt in (("fp32", torch.float32), ("fp16", torch.float16)):
        for steps, eta in ((7, 0.0), (12, 0.0), (20, 0.0), (11, 0.6)):
            shape, seeds = (2, 4, 8, 8), [420420420, 420420421]
            ref_den = kext.DiscreteEpsDDPMDenoiser(toy_eps, acp, quantize=True)
            sig_full = ks.append_zero(ref_den.t_to_sigma(torch.linspace(len(ref_den.sigmas) - 1, 0, steps)))
            sig = sig_full.to(ldt)                                   # `sigmas[...].to(self.dtype)` (:560)
            s_min, s_max = sig[sig > 0].min(), sig.max()             # 0-dim tensors in the latent dtype (:562-563)
            gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]
            x0 = (osamp.batched_randn(shape, gens, "cpu", ldt) * sig_full[0]).float()
            ns = lambda *_: osamp.batched_randn(shape, gens, "cpu", ldt).float()
            ref = ks.sample_dpm_fast(ref_den, x0, s_min, s_max, steps, eta=eta, noise_sampler=ns, disable=True)
            gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]
            den = osamp.EpsDenoiser(toy_eps, acp)
            y0 = (osamp.batched_randn(shape, gens, "cpu", ldt) * sig_full[0]).float()
            got = osamp.sample_dpm_fast(den, y0, s_min, s_max, steps,
                                        lambda *_: osamp.batched_randn(shape, gens, "cpu", ldt).float(), eta=eta)
            assert (got - ref).abs().max().item() == 0.0, f"dpm_fast {steps} {dtype_name}: oracle != vendored"
            out[f"dpm_fast/{steps}/{dtype_name}" + (f"/eta{eta}" if eta else "")] = {
                "seeds": seeds, "shape": shape, "sigmas": sig_full, "result": ref, "eta": eta, "steps": steps}
    # DPM-Solver adaptive (PID step-size control; the number of evaluations is data dependent)
    for dtype_name, ldt in (("fp32", torch.float32), ("fp16", torch.float16)):
        for steps, eta in ((10, 0.0), (25, 0.0), (10, 0.5)):
            shape, seeds = (2, 4, 8, 8), [420420420, 420420421]
            ref_den = kext.DiscreteEpsDDPMDenoiser(toy_eps, acp, quantize=True)
            sig_full = ks.append_zero(ref_den.t_to_sigma(torch.linspace(len(ref_den.sigmas) - 1, 0, steps)))
            sig = sig_full.to(ldt)
            s_min, s_max = sig[sig > 0].min(), sig.max()
            gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]
            x0 = (osamp.batched_randn(shape, gens, "cpu", ldt) * sig_full[0]).float()
            ns = lambda *_: osamp.batched_randn(shape, gens, "cpu", ldt).float()
            ref, info = ks.sample_dpm_adaptive(ref_den, x0, s_min, s_max, eta=eta, noise_sampler=ns, disable=True,
                                               return_info=True)
            gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]
            den = osamp.EpsDenoiser(toy_eps, acp)
            y0 = (osamp.batched_randn(shape, gens, "cpu", ldt) * sig_full[0]).float()
            info2 = {}
            got = osamp.sample_dpm_adaptive(den, y0, s_min, s_max,
                                            lambda *_: osamp.batched_randn(shape, gens, "cpu", ldt).float(), eta=eta,
                                            info_out=info2)
            assert info2 == info, f"dpm_adaptive: {info2} != {info}"
            assert (got - ref).abs().max().item() == 0.0, f"dpm_adaptive {steps} {dtype_name}: oracle != vendored"
            out[f"dpm_adaptive/{steps}/{dtype_name}" + (f"/eta{eta}" if eta else "")] = {
                "seeds": seeds, "shape": shape, "sigmas": sig_full, "result": ref, "eta": eta, "steps": steps, "info": info}
    # churn > 0 (Karras stochasticity) for the samplers that take s_churn
    for name in ("euler", "heun", "dpm_2"):
        shape, seeds, steps = (2, 4, 8, 8), [420420420, 420420421], 12
        gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]
        ref_den = kext.DiscreteEpsDDPMDenoiser(toy_eps, acp, quantize=True)
        sig = ks.append_zero(ref_den.t_to_sigma(torch.linspace(len(ref_den.sigmas) - 1, 0, steps)))
        x0 = (osamp.batched_randn(shape, gens, "cpu", torch.float32) * sig[0])
        ns = lambda *_: osamp.batched_randn(shape, gens, "cpu", torch.float32)

        class _T:
            def __getattr__(self, k):
                return getattr(torch, k)

            def randn_like(self, inp, **kw):
                return ns()
        ks.torch = _T()
        try:
            fn = {"euler": ks.sample_euler, "heun": ks.sample_heun, "dpm_2": ks.sample_dpm_2}[name]
            ref = fn(ref_den, x0, sig, disable=True, s_churn=6.0, s_tmin=0.5, s_tmax=9.0)
        finally:
            ks.torch = torch
        gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]
        den = osamp.EpsDenoiser(toy_eps, acp)
        y0 = osamp.batched_randn(shape, gens, "cpu", torch.float32) * sig[0]
        ofn = {"euler": osamp.sample_euler, "heun": osamp.sample_heun, "dpm_2": osamp.sample_dpm_2}[name]
        got = ofn(den, y0, sig, lambda x: osamp.batched_randn(shape, gens, "cpu", torch.float32), s_churn=6.0, s_tmin=0.5,
                  s_tmax=9.0)
        assert (got - ref).abs().max().item() == 0.0, f"{name} churn: oracle != vendored"
        out[f"{name}_churn/12/fp32"] = {"seeds": seeds, "shape": shape, "sigmas": sig, "result": ref,
                                         "churn": (6.0, 0.5, 9.0)}
    # v-prediction denoiser
    x = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(1))
    s = torch.tensor([3.3, 0.7])
    refv = kext.DiscreteVDDPMDenoiser(toy_eps, acp, quantize=True)(x, s)
    gotv = osamp.VDenoiser(toy_eps, acp)(x, s)
    assert torch.equal(refv, gotv)
    out["vdenoiser"] = {"x": x, "sigma": s, "result": refv}
    # karras
    kk = ks.get_sigmas_karras(11, ref_den.sigma_min, ref_den.sigma_max, rho=7.0)
    assert torch.equal(kk, osamp.get_sigmas_karras(11, den.sigma_min, den.sigma_max, 7.0))
    out["karras/11"] = kk
    # sigma_to_t on fp16-quantised sigmas
    sg = sig_full[:-1].to(torch.float16).float()
    assert torch.equal(ref_den.sigma_to_t(sg), den.sigma_to_t(sg))
    out["sigma_to_t"] = {"sigma": sg, "t": ref_den.sigma_to_t(sg)}
    torch.save(out, os.path.join(GOLD, "samplers.pt"))
    print("samplers pinned:", len(out))


def pin_ddim():
    mod = _vendored.gyre_ddim()
    out = {}
    for pred in ("epsilon", "v_prediction"):
        for eta in (0.0, 0.6):
            sch = mod.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                    clip_sample=False, set_alpha_to_one=False, steps_offset=1,
                                    prediction_type=pred)        # ckpt_utils.py:244-255
            n = 10
            sch.set_timesteps(n)
            assert torch.equal(sch.timesteps, osamp.ddim_timesteps(n))
            g = torch.Generator("cpu").manual_seed(7)
            x = torch.randn(2, 4, 8, 8, generator=g)
            xr = x.clone()
            gr = torch.Generator("cpu").manual_seed(99)
            for t in sch.timesteps:
                eps = toy_eps(xr, t)
                xr = sch.step(eps, t, xr, eta=eta, generator=gr).prev_sample
            go = torch.Generator("cpu").manual_seed(99)
            got = osamp.sample_ddim(toy_eps, x.clone(), n, osamp.sd_alphas_cumprod(), eta, go, pred)
            err = (got - xr).abs().max().item()
            assert err < 1e-6, f"ddim {pred} eta={eta}: {err}"
            out[f"{pred}/{eta}"] = {"x": x, "result": xr}
    torch.save(out, os.path.join(GOLD, "ddim.pt"))
    print("ddim pinned:", len(out))


def pin_tome():
    tm = _vendored.tome_merge()
    out = {}
    g = torch.Generator("cpu").manual_seed(3)
    for (B, N, C, r) in ((2, 64, 32, 16), (2, 64, 32, 40), (1, 256, 64, 128), (3, 30, 16, 7)):
        k = torch.randn(B, N, C, generator=g)
        v = torch.randn(B, N, C, generator=g)
        merge, _ = tm.bipartite_soft_matching(k, r, False, False)
        km, _ = tm.merge_wavg(merge, k)
        vm, _ = tm.merge_wavg(merge, v)
        plan = otome.bipartite_soft_matching_plan(k, r)
        assert torch.equal(otome.merge_mean(plan, k), km) and torch.equal(otome.merge_mean(plan, v), vm)
        out[f"{B}x{N}x{C}/r{r}"] = {"k": k, "v": v, "r": r, "k_merged": km, "v_merged": vm}
    sys.path.insert(0, os.path.join(_vendored.REF, "nonfree/ToMe"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("_tome_utils_src", os.path.join(_vendored.REF, "nonfree/ToMe/tome/utils.py"))
    src = open(spec.origin).read()
    ns = {}
    start = src.index("def parse_r")
    exec("from typing import List, Tuple, Union\n" + src[start:], ns)
    for arg in (8, (8, -1.0), (100, 0.5), [3, 2, 1]):
        assert ns["parse_r"](16, arg if not isinstance(arg, list) else list(arg)) == otome.parse_r(16, arg if not isinstance(arg, list) else list(arg))
    out["parse_r"] = {"(100,0.5)": ns["parse_r"](16, (100, 0.5))}
    torch.save(out, os.path.join(GOLD, "tome.pt"))
    print("tome pinned:", len(out))


def pin_clip():
    """The CLIP text-encoder oracle against the INSTALLED transformers CLIPTextModel (random-init, seeded)."""
    from transformers import CLIPTextConfig, CLIPTextModel
    from oracle import clip as oclip
    out = {}
    for act in ("quick_gelu", "gelu"):
        cfg = CLIPTextConfig(vocab_size=1000, hidden_size=64, intermediate_size=256, num_hidden_layers=3,
                             num_attention_heads=4, max_position_embeddings=77, hidden_act=act)
        torch.manual_seed(7)
        m = CLIPTextModel(cfg).eval()
        sd = {k: v.clone() for k, v in m.state_dict().items() if not k.endswith("position_ids")}
        g = torch.Generator().manual_seed(3)
        # make the norms / biases non-trivial
        for k in sd:
            if k.endswith("bias") or "layer_norm" in k:
                sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
        m.load_state_dict(sd, strict=False)
        ids = torch.randint(0, 1000, (2, 77), generator=g)
        with torch.no_grad():
            ref = m(ids, output_hidden_states=True)
        last, hs = oclip.clip_text_forward(sd, ids, num_layers=3, num_heads=4, hidden_act=act)
        err = (last - ref.last_hidden_state).abs().max().item()
        assert err < 2e-5, f"clip oracle != transformers ({act}): {err}"
        for a, b in zip(hs, ref.hidden_states):
            assert (a - b).abs().max().item() < 2e-5
        pen = m.text_model.final_layer_norm(ref.hidden_states[-2])
        assert (oclip.alt_layer(sd, ids, "penultimate", num_layers=3, num_heads=4, hidden_act=act) - pen).abs().max() < 2e-5
        out[act] = {"state_dict": sd, "ids": ids, "last_hidden_state": ref.last_hidden_state,
                    "penultimate": pen.detach(), "hidden_states": [t.detach() for t in ref.hidden_states]}
    torch.save(out, os.path.join(GOLD, "clip.pt"))
    print("clip pinned against transformers", __import__("transformers").__version__)



class FakeDiffusersUNet:
    """A deterministic stand-in for `UNet2DConditionModel` with the `DiffusersUNet` call signature
    (gyre/pipeline/unet/types.py:30-39): its output depends on the latents, the timestep, the ORDER of the text
    embeddings in the batch and on every extra input channel, so any mistake in CFG ordering, latent duplication,
    timestep duplication or extra-channel concatenation changes the result."""

    class _Out:
        def __init__(self, sample):
            self.sample = sample

    def __call__(self, latents, t, *, encoder_hidden_states, **kwargs):
        x = latents[:, :4].float()
        tt = torch.as_tensor(t).float().reshape(-1)
        tt = tt.expand(latents.shape[0]) if tt.numel() == 1 else tt
        e = encoder_hidden_states.float().mean(dim=(1, 2))
        out = 0.7 * torch.tanh(x) + 0.001 * tt[:, None, None, None] * x.roll(1, -1) + 0.1 * e[:, None, None, None]
        if latents.shape[1] > 4:
            w = torch.arange(1, latents.shape[1] - 3, dtype=torch.float32)[None, :, None, None]
            out = out + 0.05 * (latents[:, 4:].float() * w).sum(dim=1, keepdim=True)
        return self._Out(out.to(latents.dtype))


def pin_wrappers():
    """PINS the RNG contract and the CFG / embedding / extra-channel wrapper stack against the reference's own code
    (gyre/pipeline/randtools.py, unet/cfg.py, unet/core.py - pure torch, imported file by file) on a fake UNet, the
    way UnifiedPipeline composes them (unified_pipeline.py:2238, 2326-2337, 2412, 139-146)."""
    rt, _, rcfg, rcore = _vendored.gyre_pipeline_pure()
    from gyre_b200 import randtools as ours_rt
    out = {}
    # ---- batched_randn: one draw of (1, *shape[1:]) per generator per call, generators advance in lock-step
    for dt_name, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
        seeds = [420420420, 420420421, 7]
        shape = (3, 4, 8, 8)
        gr = [torch.Generator("cpu").manual_seed(s) for s in seeds]
        go = [torch.Generator("cpu").manual_seed(s) for s in seeds]
        gb = [torch.Generator("cpu").manual_seed(s) for s in seeds]
        for call in range(3):
            ref = rt.batched_randn(shape, gr, torch.device("cpu"), dt)
            assert torch.equal(ref, osamp.batched_randn(shape, go, "cpu", dt)), "oracle batched_randn"
            assert torch.equal(ref, ours_rt.batched_randn(shape, gb, torch.device("cpu"), dt)), "gyre_b200 batched_randn"
            out[f"randn_{dt_name}_{call}"] = ref
        # shape[0] a multiple of len(generators): the generator list is cycled (randtools.py:57)
        g2r = [torch.Generator("cpu").manual_seed(5), torch.Generator("cpu").manual_seed(6)]
        g2o = [torch.Generator("cpu").manual_seed(5), torch.Generator("cpu").manual_seed(6)]
        g2b = [torch.Generator("cpu").manual_seed(5), torch.Generator("cpu").manual_seed(6)]
        ref = rt.batched_randn((4, 4, 4, 4), g2r, torch.device("cpu"), dt)
        assert torch.equal(ref, osamp.batched_randn((4, 4, 4, 4), g2o, "cpu", dt))
        assert torch.equal(ref, ours_rt.batched_randn((4, 4, 4, 4), g2b, torch.device("cpu"), dt))
        out[f"randn_cycled_{dt_name}"] = ref
    # ---- wrapper stack
    g = torch.Generator("cpu").manual_seed(99)
    B = 3
    lat = torch.randn(B, 4, 8, 8, generator=g)
    cond = torch.randn(B, 77, 16, generator=g)
    unc = torch.randn(B, 77, 16, generator=g)
    extra = torch.randn(B, 5, 8, 8, generator=g)
    t_vec = torch.tensor([801, 801, 801])
    base = FakeDiffusersUNet()
    for dt_name, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
        for with_extra in (False, True):
            for t_name, t in (("tvec", t_vec), ("tint", 401)):
                unet = rcore.CFGUNetFromDiffusersUNet(base)
                kids = rcfg.CFGChildUnets(g=rcore.UNetWithEmbeddings(unet, cond.to(dt), "g"),
                                          u=rcore.UNetWithEmbeddings(unet, unc.to(dt), "u"),
                                          f=rcore.UNetWithEmbeddings(unet, torch.cat([unc, cond]).to(dt), "f"))
                if with_extra:       # EnhancedRunwayInpaintMode.wrap_unet (unified_pipeline.py:668-690)
                    kids = kids.wrap_all(rcore.UnetWithExtraChannels, extra.to(dt))
                par = rcfg.CFGUNet_Parallel(kids, 7.5, B)(lat.to(dt), t)
                seq = rcfg.CFGUNet_Sequential(kids, 7.5, B)(lat.to(dt), t)
                # oracle restatement
                x_in = lat.to(dt)
                ocfg = osamp.CFGParallel(base, unc.to(dt), cond.to(dt), 7.5)
                if with_extra:
                    class _X:            # the oracle's image modes concatenate inside eps_cfg (oracle/sampling.py)
                        def __call__(self, latents, tt, *, encoder_hidden_states, **kw):
                            e2 = torch.cat([extra.to(dt)] * (latents.shape[0] // B))
                            return base(torch.cat([latents, e2], dim=1), tt, encoder_hidden_states=encoder_hidden_states)
                    ocfg = osamp.CFGParallel(_X(), unc.to(dt), cond.to(dt), 7.5)
                mine = ocfg(x_in, t)
                assert torch.equal(par, mine), f"oracle CFGParallel vs reference ({dt_name}, extra={with_extra}, {t_name})"
                key = f"cfg_{dt_name}_{'extra' if with_extra else 'plain'}_{t_name}"
                out[key + "_parallel"] = par
                out[key + "_sequential"] = seq
    out["inputs"] = {"lat": lat, "cond": cond, "unc": unc, "extra": extra, "t_vec": t_vec, "t_int": 401, "scale": 7.5}
    torch.save(out, os.path.join(GOLD, "wrappers.pt"))
    print(f"wrappers: {len(out) - 1} reference vectors pinned (batched_randn bit-exact for the oracle and gyre_b200.randtools; "
          "CFGUNet_Parallel / Sequential + UNetWithEmbeddings + UnetWithExtraChannels bit-exact for the oracle)")


def toy_kunet(kind):
    """A deterministic stand-in for a `KDiffusionSchedulerUNet(latents, sigma, u) -> x0` leaf (types.py:63-67): depends on
    the pixel neighbourhood, on sigma and on u, and differs between the natural / hires (root / top) leaves."""
    a, b = (0.8, 0.05) if kind == "a" else (0.6, -0.08)

    def f(latents, sigma, u):
        s = torch.as_tensor(sigma).float().reshape(-1)
        s = s[:, None, None, None] if s.numel() > 1 else s
        x = latents.float()
        return (a * torch.tanh(x) + b * x.roll(1, -1) / (1 + s) + 0.1 * u * x.roll(1, -2)).to(latents.dtype)
    return f


def pin_hires():
    """PINS the hires-fix and graft wrappers (gyre/pipeline/unet/hires_fix.py, unet/graft.py, easing.py and the vendored
    ResizeRight) against the oracle restatement: lanczos2 resizes, scale_into placement, Easing, and whole wrapper
    calls over toy leaves with per-sample generators."""
    hf, gr, ea = _vendored.gyre_hires()
    from gyre import resize_right as rr
    from oracle import hires as oh
    out = {}
    g = torch.Generator("cpu").manual_seed(31)
    # ---- Easing.interp for both wrappers' curves (and the other published curves)
    for name in ("linear", "quad", "cubic", "quartic", "quintic", "sine", "circular", "expo"):
        for floor, start, end in ((0, 0, 0.667), (0, 0.1, 0.3), (0.2, 0.05, 0.9)):
            us = [i / 40 for i in range(41)]
            ref = [ea.Easing(floor=floor, start=start, end=end, easing=name).interp(u) for u in us]
            mine = [oh.Easing(floor=floor, start=start, end=end, easing=name).interp(u) for u in us]
            assert ref == mine, f"Easing {name}"
            out[f"easing/{name}/{floor}_{start}_{end}"] = {"u": us, "p": ref}
    # ---- resize + scale_into
    for dt_name, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
        for shape, scale in (((2, 4, 24, 24), 16 / 24), ((2, 4, 16, 16), 1.5), ((1, 4, 24, 16), 0.8), ((1, 4, 16, 16), 1.3),
                             ((1, 3, 48, 40), 0.7333), ((1, 4, 12, 16), 1.0)):
            x = torch.randn(shape, generator=g).to(dt)
            ref = rr.resize(x, scale_factors=scale, interp_method=rr.interp_methods.lanczos2, pad_mode="replicate",
                            antialiasing=False)
            assert torch.equal(ref, oh.resize_lanczos2(x, scale)), f"resize {shape} {scale}"
            out[f"resize/{dt_name}/{'x'.join(map(str, shape))}/{scale:.4f}"] = {"x": x, "scale": scale, "out": ref}
            for tshape in ((shape[0], shape[1], 16, 16), (shape[0], shape[1], 20, 12), (shape[0], shape[1], 40, 44)):
                if dt is torch.float16 and tshape[2] != 20:
                    continue            # keeps the fixture small: fp16 adds no placement logic
                ref_p = hf.scale_into(x, scale, target_shape=torch.Size(tshape))
                assert torch.equal(ref_p, oh.scale_into(x, scale, target_shape=tshape)), f"scale_into pad {shape} {tshape}"
                bg = torch.randn(tshape, generator=g).to(dt)
                ref_c = hf.scale_into(x, scale, target=bg.clone())
                assert torch.equal(ref_c, oh.scale_into(x, scale, target=bg.clone())), f"scale_into clone {shape} {tshape}"
                out[f"scale_into/{dt_name}/{'x'.join(map(str, shape))}/{scale:.4f}/{tshape[2]}x{tshape[3]}"] = {
                    "x": x, "scale": scale, "bg": bg, "pad": ref_p, "clone": ref_c}
    # ---- HiresUnetWrapper over toy leaves
    seeds = [420420420, 420420421]
    for dt_name, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
        for (h, w), nat, oos in (((24, 24), 16, 0.6), ((24, 16), 16, 0.6), ((16, 24), 16, 1.0), ((32, 24), 16, 0.3)):
            B = 2
            lat = torch.randn(2 * B, 4, h, w, generator=g).to(dt)
            sig = torch.tensor(3.7)
            gens_r = [torch.Generator("cpu").manual_seed(s) for s in seeds]
            gens_o = [torch.Generator("cpu").manual_seed(s) for s in seeds]

            class _Dbg:
                def log(self, *a, **k):
                    pass
            ref_w = hf.HiresUnetWrapper(toy_kunet("a"), toy_kunet("b"), gens_r, [nat, nat], oos, _Dbg())
            my_w = oh.HiresUnetWrapper(toy_kunet("a"), toy_kunet("b"), gens_o, [nat, nat], oos)
            calls = []
            for u in (0.0, 0.2, 0.45, 0.66, 0.8):        # consecutive calls: the generators advance between them
                ref = ref_w(lat, sig, u)
                mine = my_w(lat, sig, u)
                assert torch.equal(ref, mine), f"HiresUnetWrapper {dt_name} {h}x{w} u={u}"
                calls.append({"u": u, "out": ref})
            left = torch.randn(B, 4, nat, nat, generator=g).to(dt)
            right = torch.randn(B, 4, h, w, generator=g).to(dt)
            merged = hf.HiresUnetWrapper.merge_initial_latents(left, right)
            assert torch.equal(merged, oh.HiresUnetWrapper.merge_initial_latents(left, right))
            assert torch.equal(hf.HiresUnetWrapper.split_result(None, merged), oh.HiresUnetWrapper.split_result(None, merged))
            out[f"hires/{dt_name}/{h}x{w}/nat{nat}/oos{oos}"] = {"latents": lat, "sigma": sig, "seeds": seeds, "natural": nat,
                                                              "oos": oos, "calls": calls, "left": left, "right": right,
                                                              "merged": merged}
        # image_to_natural (pixels)
        img = torch.rand(1, 3, 96, 72, generator=g).to(dt)
        for oos in (0.6, 1.0):
            ref = hf.HiresUnetWrapper.image_to_natural(64, img, oos)
            assert torch.equal(ref, oh.HiresUnetWrapper.image_to_natural(64, img, oos))
            out[f"image_to_natural/{dt_name}/oos{oos}"] = {"image": img, "out": ref}
        # ---- GraftUnets
        lat = torch.randn(2, 4, 16, 16, generator=g).to(dt)
        for blend in ({}, {"start": 0.0, "end": 0.8, "easing": "linear"}):
            gens_r = [torch.Generator("cpu").manual_seed(s) for s in seeds]
            gens_o = [torch.Generator("cpu").manual_seed(s) for s in seeds]
            ref_w = gr.GraftUnets(toy_kunet("a"), toy_kunet("b"), gens_r, blend=blend)
            my_w = oh.GraftUnets(toy_kunet("a"), toy_kunet("b"), gens_o, blend=blend)
            calls = []
            for u in (0.0, 0.05, 0.15, 0.2, 0.28, 0.5, 0.9):
                ref = ref_w(lat, torch.tensor(2.5), u)
                assert torch.equal(ref, my_w(lat, torch.tensor(2.5), u)), f"GraftUnets {dt_name} u={u}"
                calls.append({"u": u, "out": ref})
            out[f"graft/{dt_name}/{'default' if not blend else 'linear'}"] = {"latents": lat, "sigma": torch.tensor(2.5),
                                                                              "seeds": seeds, "blend": blend, "calls": calls}
    torch.save(out, os.path.join(GOLD, "hires.pt"))
    print(f"hires: {len(out)} reference vectors pinned (Easing, lanczos2 resize, scale_into, HiresUnetWrapper, GraftUnets: "
          "oracle bit-exact against gyre/pipeline/unet/hires_fix.py, graft.py, easing.py and the vendored ResizeRight)")


class ToyTokenizer:
    """Deterministic stand-in for CLIPTokenizer (no vocabulary files on the box): lower-cased words and punctuation marks
    map to ids by a stable hash; `__call__(text).input_ids` = [bos] + ids + [eos] like the real tokenizer."""
    model_max_length = 77
    bos_token_id = 998
    eos_token_id = 999

    class _Out:
        def __init__(self, ids):
            self.input_ids = ids

    @staticmethod
    def _ids(text):
        import re
        import zlib
        return [1 + zlib.crc32(w.encode()) % 990 for w in re.findall(r"[a-z0-9]+|[^\sa-z0-9]", text.lower())]

    def __call__(self, text, max_length=None, truncation=False, **_):
        if isinstance(text, (list, tuple)):
            return self._Out([self(t, max_length, truncation).input_ids for t in text])
        ids = [self.bos_token_id] + self._ids(text) + [self.eos_token_id]
        if truncation and max_length is not None and len(ids) > max_length:
            ids = ids[:max_length - 1] + [self.eos_token_id]
        return self._Out(ids)


LPW_PROMPTS = [
    "a (very beautiful:1.3) masterpiece, [dull] colours, ((sharp)) focus",
    "an \\(escaped\\) bracket and a lone : colon (unbalanced",
    "plain prompt without any weighting at all",
    " ".join(f"(word{i}:{1 + (i % 7) / 10:.1f}) filler{i}," for i in range(60)),          # > 150 tokens: three chunks
    "",
]
LPW_NEGATIVE = ["blurry, (low quality:1.4), [[watermark]]", "", "text", "(bad:1.2) " * 50, "ugly"]


def pin_lpw():
    """PINS the LPW prompt-weighting front end against the reference's own lpw_text_embedding.py: the bracket grammar,
    token / weight lists, padding, chunked encoding and the mean-preserving weighting, with the installed transformers
    CLIPTextModel (random init) as the text encoder and a toy tokenizer."""
    from transformers import CLIPTextConfig, CLIPTextModel
    from gyre_b200 import lpw_text_embedding as mine
    ref = _vendored.gyre_lpw()
    out = {}
    # ---- the grammar, on the reference's own doctest strings and more
    cases = ["normal text", "an (important) word", "(unbalanced", "\\(literal\\]", "(unnecessary)(parens)",
             "a (((house:1.3)) [on] a (hill:0.5), sun, (((sky))).", "a:b (c:d) e:1.2) [f:2.0] (g:-0.5) (h:+.5)", "\\", "a\\b",
             "((a:1.2):0.5) [b (c] d)", ":", "(:1.1)", "x (y:1.) z", "]) stray closers [("] + LPW_PROMPTS + LPW_NEGATIVE
    parsed = []
    for c in cases:
        r = ref.parse_prompt_attention(c)
        assert mine.parse_prompt_attention(c) == r, f"parse_prompt_attention({c!r}): {mine.parse_prompt_attention(c)} vs {r}"
        parsed.append(r)
    out["parse"] = {"cases": cases, "parsed": parsed}
    tok = ToyTokenizer()
    cfg = CLIPTextConfig(vocab_size=1000, hidden_size=64, intermediate_size=256, num_hidden_layers=3,
                         num_attention_heads=4, max_position_embeddings=77, hidden_act="quick_gelu")
    torch.manual_seed(11)
    model = CLIPTextModel(cfg).eval()
    g = torch.Generator().manual_seed(5)
    sd = {}
    for k, v in model.state_dict().items():
        if k.endswith("position_ids"):
            continue
        if "layer_norm" in k:      # non-trivial affine parameters
            v = v + 0.2 * torch.randn(v.shape, generator=g)
        sd[k] = v.half().float()
    model.load_state_dict(sd, strict=False)
    out["clip_config"] = dict(vocab_size=1000, hidden_size=64, intermediate_size=256, num_hidden_layers=3,
                              num_attention_heads=4, max_position_embeddings=77, hidden_act="quick_gelu")
    out["clip_state_dict"] = {k: v.half() for k, v in sd.items()}
    with torch.no_grad():
        for mult in (1, 3):
            for bos_mid in (False, True):
                max_len = (tok.model_max_length - 2) * mult + 2
                t_ref, w_ref = ref.get_prompts_with_weights(tok, LPW_PROMPTS, max_len - 2)
                t_mine, w_mine = mine.get_prompts_with_weights(tok, LPW_PROMPTS, max_len - 2)
                assert t_ref == t_mine and w_ref == w_mine, "get_prompts_with_weights"
                emb, unc = ref.get_weighted_text_embeddings(tok, model, model, torch.device("cpu"), list(LPW_PROMPTS),
                                                            list(LPW_NEGATIVE), max_embeddings_multiples=mult,
                                                            no_boseos_middle=bos_mid)
                raw, _ = ref.get_weighted_text_embeddings(tok, model, model, torch.device("cpu"), list(LPW_PROMPTS), None,
                                                          max_embeddings_multiples=mult, no_boseos_middle=bos_mid,
                                                          skip_weighting=True)
                # padded tokens / weights as the reference builds them (re-derived here for the host-logic test)
                longest = max(max(len(t) for t in t_ref), max(len(t) for t in ref.get_prompts_with_weights(
                    tok, LPW_NEGATIVE, max_len - 2)[0]))
                m2 = max(1, min(mult, (longest - 1) // 75 + 1))
                ml = 75 * m2 + 2
                pt, pw = ref.pad_tokens_and_weights([list(t) for t in t_ref], [list(w) for w in w_ref], ml, tok.bos_token_id,
                                                    tok.eos_token_id, no_boseos_middle=bos_mid, chunk_length=77)
                pt2, pw2 = mine.pad_tokens_and_weights([list(t) for t in t_ref], [list(w) for w in w_ref], ml,
                                                       tok.bos_token_id, tok.eos_token_id, no_boseos_middle=bos_mid,
                                                       chunk_length=77)
                assert pt == pt2 and pw == pw2, "pad_tokens_and_weights"
                out[f"lpw/mult{mult}/{'nomid' if bos_mid else 'mid'}"] = {
                    "tokens": torch.tensor(pt), "weights": torch.tensor(pw), "text": emb, "uncond": unc}
                if mult == 3 and not bos_mid:
                    out[f"lpw/mult{mult}/mid"]["unweighted"] = raw
    torch.save(out, os.path.join(GOLD, "lpw.pt"))
    print(f"lpw: grammar ({len(cases)} strings), token / weight lists and padding identical to the reference's; "
          f"{sum(k.startswith('lpw/') for k in out)} weighted-embedding vectors stored")


def pin_images():
    """PINS the outpaint image tail against the reference's own numpy histogram matching (gyre/match_histograms.py, pure
    numpy, imported as is) wrapped in the tensor statements of unified_pipeline.py:2493-2510 and the uint8 conversions of
    gyre/images.py:64-83, 667-672 (images.py itself imports cv2 / PIL at module level: its four conversion lines are
    restated here, cited)."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("_gyre_match_histograms", os.path.join(_vendored.REF, "gyre/match_histograms.py"))
    mh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mh)

    def to_cv(t):        # images.py:69-83 (the BGR flip permutes channels consistently on both sides: omitted)
        return (t.permute(0, 2, 3, 1).to(torch.float32) * 255).round().to(torch.uint8).cpu().numpy()

    def from_cv(a):      # images.py:50-66
        return (torch.from_numpy(a).to(torch.float32) / 255.0).permute(0, 3, 1, 2)

    out = {}
    g = torch.Generator().manual_seed(41)
    for name, (B, H, W) in (("small", (2, 24, 40)), ("batch3", (3, 64, 64))):
        for dt_name, dt in (("fp16", torch.float16),):
            result = torch.rand(B, 3, H, W, generator=g).pow(1.7).to(dt)                 # skewed histogram
            source = (torch.rand(1, 3, H, W, generator=g) * 0.8 + 0.1).to(dt).expand(B, -1, -1, -1).contiguous()
            mask = torch.zeros(1, 3, H, W)
            mask[:, :, H // 4:, W // 3:] = 1.0
            mask[:, :, H // 4:H // 2, W // 3:W // 2] = 0.5                               # a soft edge
            mask = mask.to(dt).expand(B, -1, -1, -1).contiguous()
            reference = source * (1 - mask) + result * mask
            matched = mh.match_histograms(to_cv(result), to_cv(reference), channel_axis=3)
            assert matched.dtype == np.uint8
            res = from_cv(matched).to(result)
            final = source * (1 - mask) + res * mask
            out[f"{name}/{dt_name}"] = {"result": result, "source": source, "outmask": mask, "final": final,
                                        "matched_u8": torch.from_numpy(matched)}
    torch.save(out, os.path.join(GOLD, "images.pt"))
    print(f"images: {len(out)} outpaint histogram-match vectors from gyre/match_histograms.py stored")


def pin_t2i_adapter():
    """PINS the T2I-adapter encoder oracle against the reference's own gyre/pipeline/t2i_adapter/adapter.py (`Adapter`)."""
    from oracle import t2i_adapter as oad
    ad = _vendored.gyre_t2i_adapter()
    out = {}
    for name, kw in (("main_tiny", dict(channels=[32, 64, 96, 96], nums_rb=2, cin=192, ksize=1, sk=True, use_conv=False)),
                     ("conv_tiny", dict(channels=[32, 32, 32], nums_rb=2, cin=64, ksize=3, sk=False, use_conv=True))):
        torch.manual_seed(17)
        m = ad.Adapter(**kw).eval()
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        shapes = oad.adapter_param_shapes(**kw)
        assert {k: tuple(v.shape) for k, v in sd.items()} == shapes, f"adapter parameter inventory ({name})"
        g = torch.Generator().manual_seed(3)
        cin_img = kw["cin"] // 64
        x = torch.rand(2, cin_img, 128, 96, generator=g)
        with torch.no_grad():
            ref = m(x)
            mine = oad.adapter_forward(sd, x, **{k: v for k, v in kw.items() if k != "cin"})
        for a_, b_ in zip(ref, mine):
            assert torch.equal(a_, b_), f"Adapter.forward ({name})"
        # (sk=False only works upstream when consecutive levels have equal widths: `skep` is declared on in_c channels but
        # applied after in_conv, adapter.py:76-78, 91-97 - hence the constant widths of the second configuration)
        out[name] = {"config": kw, "state_dict": {k: v.half() for k, v in sd.items()}, "x": x.half()}
        # features of the fp16-rounded weights / input, evaluated in fp32 (what the native fp16 path is compared with)
        sd16 = {k: v.half().float() for k, v in sd.items()}
        with torch.no_grad():
            out[name]["features"] = oad.adapter_forward(sd16, x.half().float(), **{k: v for k, v in kw.items() if k != "cin"})
    # Adapter_light (`type: light`)
    lkw = dict(channels=[32, 64, 96, 96], nums_rb=2, cin=192)
    torch.manual_seed(19)
    m = ad.Adapter_light(**lkw).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    assert {k: tuple(v.shape) for k, v in sd.items()} == oad.adapter_light_param_shapes(**lkw), "light adapter parameter inventory"
    x = torch.rand(2, 3, 128, 96, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = m(x)
        mine = oad.adapter_light_forward(sd, x, channels=lkw["channels"], nums_rb=lkw["nums_rb"])
        for a_, b_ in zip(ref, mine):
            assert torch.equal(a_, b_), "Adapter_light.forward"
        sd16 = {k: v.half().float() for k, v in sd.items()}
        out["light_tiny"] = {"config": lkw, "state_dict": {k: v.half() for k, v in sd.items()}, "x": x.half(),
                             "features": oad.adapter_light_forward(sd16, x.half().float(), channels=lkw["channels"],
                                                                   nums_rb=lkw["nums_rb"])}
    # StyleAdapter (`type: style`)
    skw = dict(width=64, context_dim=48, num_head=4, n_layes=2, num_token=4)
    torch.manual_seed(23)
    m = ad.StyleAdapter(**skw).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g3 = torch.Generator().manual_seed(6)
    for k in sd:
        if k.endswith("bias") or "ln_" in k:
            sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g3)
    m.load_state_dict(sd)
    xs = torch.randn(2, 17, skw["width"], generator=g3)
    with torch.no_grad():
        ref = m(xs)
        mine = oad.style_adapter_forward(sd, xs, num_head=skw["num_head"], num_token=skw["num_token"])
        err = (ref - mine).abs().max().item()
        assert err < 2e-6, f"StyleAdapter.forward: {err}"
        sd16 = {k: v.half().float() for k, v in sd.items()}
        out["style_tiny"] = {"config": skw, "state_dict": {k: v.half() for k, v in sd.items()}, "x": xs.half(),
                             "tokens": oad.style_adapter_forward(sd16, xs.half().float(), num_head=skw["num_head"],
                                                                 num_token=skw["num_token"])}
    torch.save(out, os.path.join(GOLD, "t2i_adapter.pt"))
    print(f"t2i_adapter: {len(out)} configurations, oracle bit-exact against gyre/pipeline/t2i_adapter/adapter.py")


def pin_safety():
    """PINS the safety-checker oracle: the resize against Pillow's own `Image.resize(BICUBIC)` (bit-exact), the vision
    tower / projection / cosine scores / flag loop against the reference's FlagOnlySafetyChecker
    (gyre/pipeline/safety_checkers.py) built on the installed transformers.  (CLIPFeatureExtractor 4.28's rescale /
    normalise lines are restated - transformers 5.5's processor resizes with torch and lands one LSB away from Pillow,
    so it is not the pin target.)"""
    import numpy as np
    from PIL import Image
    from transformers import CLIPConfig
    from oracle import safety as osf
    sc = _vendored.gyre_safety_checkers()
    out = {"resize": [], "models": {}}
    for (h, w) in [(512, 512), (64, 96), (300, 224), (160, 96), (100, 37), (224, 224)]:
        img = osf.synthetic_image(h, w)
        nh, nw = osf.resize_output_size(h, w, 224)
        ref = np.asarray(Image.fromarray(img).resize((nw, nh), resample=Image.BICUBIC))
        got = osf.pil_resize_bicubic(img, nw, nh)
        assert np.array_equal(ref, got), f"PIL resize {(h, w)} -> {(nh, nw)}"
        pv = osf.clip_preprocess(img[None])
        # the crop / rescale / normalise lines on Pillow's own resize output
        top, left = (nh - 224) // 2, (nw - 224) // 2
        x = (ref[top:top + 224, left:left + 224] * (1 / 255)).astype(np.float32)
        x = ((x - np.array(osf.CLIP_MEAN, np.float32)) / np.array(osf.CLIP_STD, np.float32)).transpose(2, 0, 1)
        assert np.array_equal(pv[0], x)
        keep_resized = (h, w) in ((64, 96), (160, 96))
        out["resize"].append({"image_hw": (h, w), "size": (nh, nw),
                              "resized": torch.from_numpy(ref.copy()) if keep_resized else None,
                              "resized_sum": int(ref.astype(np.int64).sum()),
                              "resized_crc": int(np.bitwise_xor.reduce((ref.astype(np.int64).ravel() * (np.arange(ref.size) % 65521 + 1)) % (1 << 31))),
                              "pixel_values_f16": torch.from_numpy(pv[0]).half()})
    for name, vis, proj in (("tiny", dict(image_size=56, patch_size=14, hidden_size=64, intermediate_size=256, num_hidden_layers=2,
                                         num_attention_heads=4, hidden_act="quick_gelu"), 32),
                            ("vit224", dict(image_size=224, patch_size=14, hidden_size=128, intermediate_size=512,
                                            num_hidden_layers=2, num_attention_heads=2, hidden_act="quick_gelu"), 64),
                            ("gelu", dict(image_size=64, patch_size=16, hidden_size=64, intermediate_size=128, num_hidden_layers=1,
                                          num_attention_heads=1, hidden_act="gelu"), 40)):
        cfg = CLIPConfig(vision_config=vis, projection_dim=proj)
        torch.manual_seed(23)
        m = sc.FlagOnlySafetyChecker(cfg).eval()
        g = torch.Generator().manual_seed(9)
        sd = {k: v.clone() for k, v in m.state_dict().items() if not k.endswith("position_ids")}
        for k in sd:
            if k.endswith("bias") or "norm" in k:
                sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
        sd["vision_model.vision_model.embeddings.class_embedding"] = torch.randn(vis["hidden_size"], generator=g) * 0.5
        # random-init CLIP barely looks at its input (patch weights ~0.02 against unit position embeddings): make it look
        pe = "vision_model.vision_model.embeddings.patch_embedding.weight"
        sd[pe] = sd[pe] * (1.0 / sd[pe].std()) * 0.08
        for k in sd:
            if k.endswith(("v_proj.weight", "out_proj.weight", "fc2.weight")):
                sd[k] = sd[k] * 3.0
        sd["concept_embeds"] = torch.randn(17, proj, generator=g)
        sd["special_care_embeds"] = torch.randn(3, proj, generator=g)
        B = 4 if name == "vit224" else 6
        S = vis["image_size"]
        x = (torch.randn(B, 3, S, S, generator=g) * torch.linspace(0.5, 2.5, B)[:, None, None, None]
             + torch.randn(B, 3, 1, 1, generator=g)).half().float()
        # everything the native path rounds to fp16 on load is rounded here too, then evaluated in fp32
        sd = {k: (v if k.endswith("embeds") or k.endswith("embeds_weights") or "norm" in k or k.endswith("bias")
                  or k.endswith("class_embedding") else v.half().float()) for k, v in sd.items()}
        m.load_state_dict(sd, strict=False)
        P = {k[len("vision_model."):] if k.startswith("vision_model.vision_model.") else k: v for k, v in sd.items()}
        with torch.no_grad():
            vout = m.vision_model(x, output_hidden_states=True, return_dict=True)
            pooled_ref = vout.pooler_output
            emb_ref = m.visual_projection(pooled_ref)
            pooled, emb, hs = osf.clip_vision_forward(P, x, num_layers=vis["num_hidden_layers"], num_heads=vis["num_attention_heads"],
                                                      patch_size=vis["patch_size"], hidden_act=vis["hidden_act"],
                                                      return_hidden_states=True)
            assert (pooled - pooled_ref).abs().max().item() < 2e-5 and (emb - emb_ref).abs().max().item() < 2e-5, name
            assert len(hs) == len(vout.hidden_states)
            for a_, b_ in zip(hs, vout.hidden_states):
                assert (a_ - b_).abs().max().item() < 5e-5, name          # hidden states (the style adapter's input)
            assert (hs[-1] - vout.last_hidden_state).abs().max().item() < 5e-5
            scores = osf.cosine_scores(emb_ref, P)
        print(f"  {name}: score spread across the batch (std per column, mean) {scores.std(0).mean().item():.3f}")
        # thresholds in the middle of the score distribution so the flags are mixed, special-care hits included, and no
        # margin closer to zero than the fp16 tower's error
        def gap_threshold(col, high):
            v = col.sort().values
            gaps = v[1:] - v[:-1]
            ok = [i for i in range(len(gaps)) if gaps[i] > 0.05 and (i >= len(gaps) // 2 if high else True)] or [int(gaps.argmax())]
            i = ok[int(torch.randint(len(ok), (1,), generator=g))]
            return (v[i] + v[i + 1]) / 2

        for attempt in range(200):
            # most concepts never fire (threshold above every score); a few columns split the batch at a gap
            thr_s = scores[:, :3].max(0).values + 0.05
            thr_c = scores[:, 3:].max(0).values + 0.05
            i = int(torch.randint(3, (1,), generator=g))
            thr_s[i] = gap_threshold(scores[:, i], True)
            for j in torch.randperm(17, generator=g)[:3].tolist():
                thr_c[j] = gap_threshold(scores[:, 3 + j], True)
            res, flags = osf.flag_only(scores.numpy(), thr_s, thr_c)
            margins = [abs(v) for r in res for v in list(r["special_scores"].values()) + list(r["concept_scores"].values())]
            n_special = sum(len(r["special_care"]) > 0 for r in res)
            if min(margins) > 0.008 and 0 < sum(flags) < B and 0 < n_special < B:
                break
        else:
            raise AssertionError(f"no robust thresholds found ({name}): min margin {min(margins)}, flags {flags}, special {n_special}\n{scores}")
        sd["special_care_embeds_weights"], sd["concept_embeds_weights"] = thr_s.clone(), thr_c.clone()
        m.load_state_dict(sd, strict=False)
        imgs = np.zeros((B, 8, 8, 3), np.float32)
        with torch.no_grad():
            imgs_out, flags_ref = m(clip_input=x, images=imgs)
        assert imgs_out is imgs and list(flags_ref) == list(flags), f"FlagOnlySafetyChecker.forward ({name}): {flags_ref} vs {flags}"
        sd = {k: (v.half() if torch.equal(v.half().float(), v) else v) for k, v in sd.items()}
        out["models"][name] = {"vision_config": vis, "projection_dim": proj, "state_dict": sd, "clip_input": x.half(),
                               "hidden_last": hs[-1][:2].half() if name == "tiny" else None,
                               "hidden_penultimate": hs[-2][:2].half() if name == "tiny" else None,
                               "image_embeds": emb_ref, "scores": scores, "flags": [bool(f) for f in flags],
                               "result": [{"special_scores": [float(v) for v in r["special_scores"].values()],
                                           "concept_scores": [float(v) for v in r["concept_scores"].values()],
                                           "bad_concepts": [int(v) for v in r["bad_concepts"]]} for r in res]}
    torch.save(out, os.path.join(GOLD, "safety.pt"))
    print(f"safety: {len(out['resize'])} resizes bit-exact against Pillow {__import__('PIL').__version__}; "
          f"{len(out['models'])} checkers against gyre/pipeline/safety_checkers.py on transformers {__import__('transformers').__version__}")


def pin_controlnet():
    """PINS what gyre/pipeline/controlnet/models.py itself states - ControlNetConditioningEmbedding, the zero-convolution
    list, the parameter inventory of those parts, timestep handling, channel-order flip and the forward wiring - against
    oracle/controlnet.py, by running the reference's ControlNetModel with the absent diffusers blocks replaced by stand-ins
    that evaluate the oracle's own blocks (scripts/_vendored.py:gyre_controlnet).  The blocks themselves stay unpinned."""
    from oracle import controlnet as ocn
    cn = _vendored.gyre_controlnet()
    cfg = UNetConfig.tiny()
    out = {}
    for name, order in (("rgb", "rgb"), ("bgr", "bgr")):
        torch.manual_seed(29)
        m = cn.ControlNetModel(in_channels=cfg.in_channels, block_out_channels=cfg.block_out_channels,
                               layers_per_block=cfg.layers_per_block, cross_attention_dim=cfg.cross_attention_dim,
                               attention_head_dim=cfg.num_heads[0], norm_num_groups=cfg.norm_num_groups, norm_eps=cfg.norm_eps,
                               use_linear_projection=cfg.use_linear_projection,
                               controlnet_conditioning_channel_order=order).eval()
        g = torch.Generator().manual_seed(13)
        # (the reference zero-initialises the 14 output convolutions and the embedding's conv_out: seeded values everywhere)
        sd = synth_params(ocn.controlnet_param_shapes(cfg), seed=77)
        assert set(sd) == set(m.state_dict())
        m.load_state_dict(sd)
        shapes = ocn.controlnet_param_shapes(cfg)
        assert {k: tuple(v.shape) for k, v in sd.items()} == shapes, "ControlNet parameter inventory"
        x = torch.randn(2, cfg.in_channels, 16, 16, generator=g)
        ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
        cond = torch.rand(2, 3, 128, 128, generator=g).half().float()
        cases = {"tvec": torch.tensor([981, 21]), "tscalar": torch.tensor(500), "tint": 37}
        res = {}
        for tn, t in cases.items():
            with torch.no_grad():
                ref = m(x.clone(), t, ctx, cond, return_dict=False)
                cond_o = torch.flip(cond, dims=[1]) if order == "bgr" else cond
                mine = ocn.controlnet_forward(sd, cfg, x, t, ctx, cond_o)
            assert len(ref[0]) == len(mine[0]) == 12
            for a_, b_ in zip(ref[0] + (ref[1],), mine[0] + (mine[1],)):
                assert torch.equal(a_, b_), f"ControlNetModel.forward ({name}, {tn})"
            res[tn] = {"down": [t_.clone() for t_ in ref[0]], "mid": ref[1].clone()}
        with torch.no_grad():
            e_ref = m.controlnet_cond_embedding(cond)
            assert torch.equal(e_ref, ocn.cond_embedding(sd, cond))
        out[name] = {"weights_seed": 77, "x": x, "ctx": ctx, "cond": cond.half(), "t": {k: v for k, v in cases.items()},
                     # per-tensor sums: the full tensors are regenerated by the oracle in the tests, these pin them
                     "sums": {tn: [float(t_.double().sum()) for t_ in r["down"]] + [float(r["mid"].double().sum())]
                              for tn, r in res.items()}}
    torch.save(out, os.path.join(GOLD, "controlnet.pt"))
    print("controlnet: ControlNetModel.forward (rgb / bgr, 3 timestep forms), conditioning embedding and parameter inventory "
          "bit-exact against gyre/pipeline/controlnet/models.py (diffusers blocks stood in for by the oracle's)")


def pin_hints():
    """PINS the hint wrappers (oracle/hints.py UNetWithControlnet / AdapterStateList / UNetWithT2I, and the product's copies in
    gyre_b200/hints.py, which are plain torch) against the reference's own gyre/pipeline/unet/core.py classes over fake
    ControlNets / adapters / UNet whose outputs depend on every argument they are handed."""
    from types import SimpleNamespace as SN
    from oracle import hints as oh
    _, _, _, core = _vendored.gyre_pipeline_pure()
    g = torch.Generator().manual_seed(41)

    from fakes import FakeHintAdapter as FakeAdapter, FakeHintControlnet as FakeControlnet, FakeHintUNet as FakeUNet
    from fakes import FakeHintStyleAdapter as FakeStyle
    lat = torch.randn(2, 4, 8, 8, generator=g)
    t = torch.tensor([500, 500])
    ehs = torch.randn(2, 5, 6, generator=g)
    out = {"inputs": {"lat": lat, "t": t, "ehs": ehs}, "controlnet": {}, "t2i": {}}
    for meta in ("f", "g", "u"):
        for combo in ((False,), (True,), (False, True)):
            cns = [FakeControlnet(100 * (i + 1), c) for i, c in enumerate(combo)]
            if meta == "u" and all(combo):
                continue          # the reference hands sum(ints) to the UNet there; nothing to compare
            ref = core.UNetWithControlnet(FakeUNet(), cns)(lat, t, encoder_hidden_states=ehs, cfg_meta=meta)
            mine = oh.UNetWithControlnet(FakeUNet(), cns)(lat, t, encoder_hidden_states=ehs, cfg_meta=meta)
            assert torch.equal(ref, mine), ("UNetWithControlnet", meta, combo)
            out["controlnet"][f"{meta}/{''.join('c' if c else 'b' for c in combo)}"] = ref
        for combo in ((False,), (True,), (False, True), (True, True)):
            ads = [FakeAdapter(1000 * (i + 1), c) for i, c in enumerate(combo)]
            e = torch.cat([ehs[:1], ehs[:1]]) if meta == "f" else ehs[:1]
            l = torch.cat([lat[:1], lat[:1]]) if meta == "f" else lat[:1]
            tt = t[:2] if meta == "f" else t[:1]
            ref_w = core.UNetWithT2I(FakeUNet(), ads)
            mine_w = oh.UNetWithT2I(FakeUNet(), ads)
            ref = ref_w(l, tt, encoder_hidden_states=e, cfg_meta=meta)
            assert torch.equal(ref, mine_w(l, tt, encoder_hidden_states=e, cfg_meta=meta)), ("UNetWithT2I", meta, combo)
            # cfg_meta inferred from the embedding batch when absent (core.py:213-214)
            assert torch.equal(ref_w(l, tt, encoder_hidden_states=e), mine_w(l, tt, encoder_hidden_states=e))
            for k in ("u", "g", "f"):
                for a_, b_ in zip(ref_w.standard_states[k], mine_w.standard_states[k]):
                    assert torch.equal(a_, b_)
            out["t2i"][f"{meta}/{''.join('c' if c else 'b' for c in combo)}"] = ref
    # style adapters (context tokens for the guided side, core.py:221-237) next to a standard one
    out["t2i_style"] = {}
    for meta in ("f", "g", "u"):
        for n_style in (1, 2):
            ads = [FakeAdapter(1000, False)] + [FakeStyle(50 + i, tokens=2) for i in range(n_style)]
            e = torch.cat([ehs[:1], ehs[1:2]]) if meta == "f" else ehs[:1]
            l = torch.cat([lat[:1], lat[:1]]) if meta == "f" else lat[:1]
            tt = t[:2] if meta == "f" else t[:1]

            class CtxUNet(FakeUNet):          # the context must matter token by token here
                def __call__(self, latents, t_, **kw):
                    w = torch.arange(1, kw["encoder_hidden_states"].shape[1] + 1, dtype=torch.float32)[None, :, None]
                    return super().__call__(latents, t_, **kw) + (kw["encoder_hidden_states"] * w).mean(dim=(1, 2))[:, None, None, None]
            ref = core.UNetWithT2I(CtxUNet(), ads)(l, tt, encoder_hidden_states=e, cfg_meta=meta)
            mine = oh.UNetWithT2I(CtxUNet(), ads)(l, tt, encoder_hidden_states=e, cfg_meta=meta)
            assert torch.equal(ref, mine), ("UNetWithT2I style", meta, n_style)
            out["t2i_style"][f"{meta}/{n_style}"] = ref
    rl, ml = core.AdapterStateList(), oh.AdapterStateList()
    for i, c in enumerate((False, True, False)):
        st = FakeAdapter(7 + i, c).state
        rl.append(st, c)
        ml.append(st, c)
    for prop in ("all", "cfg_only", "either"):
        for a_, b_ in zip(getattr(rl, prop), getattr(ml, prop)):
            assert all(torch.equal(x_, y_) for x_, y_ in zip(a_, b_)), prop
    torch.save(out, os.path.join(GOLD, "hints.pt"))
    print(f"hints: UNetWithControlnet ({len(out['controlnet'])} cases), UNetWithT2I ({len(out['t2i'])} + {len(out['t2i_style'])} style cases), AdapterStateList "
          "bit-exact against gyre/pipeline/unet/core.py")


def pin_attention():
    """PINS the oracle's attention module (oracle/unet.py:attention - projections, head split, softmax(QK^T/sqrt d)V, output
    projection, and the ToMe K/V merge inside it) against the reference's own modules: MemoryEfficientCrossAttention
    (gyre/pipeline/models/memory_efficient_cross_attention.py) and ToMeMemoryEfficientCrossAttention
    (nonfree/tome_memory_efficient_cross_attention.py, with the REAL vendored tome.merge).  Stand-ins: xformers'
    memory_efficient_attention by its definition, diffusers' CrossAttention base class by its attribute set."""
    from oracle.unet import attention as oattn
    mea, tmea = _vendored.gyre_attention_modules()
    out = {}
    g = torch.Generator().manual_seed(19)
    for name, (C_, heads, N_, ctx_dim, L, r) in {"self": (64, 4, 64, None, 0, 0), "cross": (64, 4, 64, 48, 77, 0),
                                                  "self_d40": (80, 2, 96, None, 0, 0),
                                                  "tome_r16": (64, 4, 64, None, 0, 16), "tome_half": (64, 4, 64, None, 0, 32),
                                                  "tome_odd": (64, 2, 51, None, 0, 20)}.items():
        torch.manual_seed(3)
        d = C_ // heads
        if r:
            m = tmea.ToMeMemoryEfficientCrossAttention(C_, ctx_dim, heads=heads, dim_head=d).eval()
            m._tome_info = {"r": [r], "class_token": False, "distill_token": False, "trace_source": False, "size": None,
                            "source": None}
        else:
            m = mea.MemoryEfficientCrossAttention(C_, ctx_dim, heads=heads, dim_head=d).eval()
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        sd["to_out.0.bias"] = 0.1 * torch.randn(C_, generator=g)
        m.load_state_dict(sd)
        x = torch.randn(2, N_, C_, generator=g)
        if r:
            # ToMe: tokens whose merge plan survives fp16 arithmetic (the GPU test runs the same module in fp16) - every even
            # token is a noisy copy of one odd token, the noise levels spread so that the scores that rank the merges differ
            # by more than fp16 resolves; the plan is checked below to have no near-ties on the projected keys
            for attempt in range(200):
                nb = N_ // 2
                base = torch.randn(2, nb, C_, generator=g)
                na = N_ - nb
                cos = torch.linspace(0.97, 0.55, na)
                noise = torch.randn(2, na, C_, generator=g)
                src = base[:, torch.randperm(nb, generator=g)[torch.arange(na) % nb]]
                a_tok = src + noise * (1 / cos ** 2 - 1).sqrt()[None, :, None]
                x = torch.empty(2, N_, C_)
                x[:, 0::2], x[:, 1::2] = a_tok, base
                k_ = torch.nn.functional.linear(x, sd["to_k.weight"])
                k_ = k_ / k_.norm(dim=-1, keepdim=True)
                sc = k_[:, 0::2] @ k_[:, 1::2].transpose(-1, -2)
                top2 = sc.topk(2, dim=-1).values
                # what has to be unambiguous: each token's best partner, and WHICH r tokens are merged (the order inside the two
                # sets does not matter - attention is invariant to the order of its keys)
                best = top2[..., 0].sort(dim=-1, descending=True).values
                r_eff = min(r, na)
                edge = (best[:, r_eff - 1] - best[:, r_eff]).min().item() if r_eff < na else 1.0
                if (top2[..., 0] - top2[..., 1]).min() > 0.02 and edge > 5e-3:
                    break
            else:
                raise AssertionError("no tie-free ToMe fixture found")
        ctx = torch.randn(2, L, ctx_dim, generator=g) if ctx_dim else None
        with torch.no_grad():
            ref = m(x, context=ctx)
            P = {f"a.{k}": v for k, v in sd.items()}
            mine = oattn(P, "a", x, ctx, heads, tome_r=r)
        err = (ref - mine).abs().max().item()
        assert err <= 2e-6, f"attention module ({name}): {err}"
        out[name] = {"config": (C_, heads, N_, ctx_dim, L, r), "state_dict": sd, "x": x, "ctx": ctx, "out": ref}
        print(f"  attention {name}: max abs diff {err:.2e}")
    torch.save(out, os.path.join(GOLD, "attention.pt"))
    print("attention: oracle module == MemoryEfficientCrossAttention / ToMeMemoryEfficientCrossAttention")


def _reference_segment(up, unet_like, vae_like, unc, emb, *, kind, sampler_fn, steps, seeds, height, width, sample_size,
                       image=None, mask_image=None, strength=0.8, guidance_scale=7.5, cfg_execution="parallel",
                       latents_dtype=torch.float32, prediction_type="epsilon", eta=None, karras_rho=None, hints=()):
    """The hot segment of UnifiedPipeline.__call__ (unified_pipeline.py:2284-2483) driven with the REFERENCE's own classes -
    CFGUNetFromDiffusersUNet / UNetWithEmbeddings / CFGChildUnets / CFGUNet_* (unet/core.py, unet/cfg.py), the mode classes
    (unified_pipeline.py:126-696), KDiffusionScheduler (common_scheduler.py) over the vendored k-diffusion - around a
    DiffusersUNet-protocol object and a VAE-protocol object.  The glue below is the call sequence of `__call__`."""
    from types import SimpleNamespace as SN
    cs = sys.modules["gyre.pipeline.common_scheduler"]
    core = sys.modules["gyre.pipeline.unet.core"]
    cfgm = sys.modules["gyre.pipeline.unet.cfg"]
    dev = torch.device("cpu")
    B = len(seeds)
    generators = [torch.Generator(device="cpu").manual_seed(s_) for s_ in seeds]
    pipeline = SN(vae_scale_factor=8, execution_device=dev, unet=unet_like, vae=vae_like, vae_dtype=torch.float32,
                  get_unet_sample_size=lambda u: sample_size)
    cscheduler = cs.KDiffusionScheduler(sampler_fn, generators, dev, latents_dtype)
    unet = core.CFGUNetFromDiffusersUNet(unet_like)
    grouped = {}
    for h in hints:
        grouped.setdefault(type(h), []).append(h)
    for cls, hs in grouped.items():
        unet = cls.wrap_unet(unet, list(hs))
    unet_cfg = cfgm.CFGChildUnets(g=core.UNetWithEmbeddings(unet, emb, "g"), u=core.UNetWithEmbeddings(unet, unc, "u"),
                                  f=core.UNetWithEmbeddings(unet, torch.cat([unc, emb]), "f"))
    common = dict(pipeline=pipeline, scheduler=cscheduler, generators=generators, width=width, height=height, image=image,
                  mask_image=mask_image, latents_dtype=latents_dtype, batch_total=B, num_inference_steps=steps,
                  strength=strength, do_classifier_free_guidance=True, cfg_execution=cfg_execution,
                  latent_debugger=SN(log=lambda *a, **k: None))
    mode = {"txt2img": up.Txt2imgMode, "img2img": up.Img2imgMode, "inpaint": up.EnhancedInpaintMode,
            "runway": up.EnhancedRunwayInpaintMode}[kind](**common)
    unet_cfg = unet_cfg.wrap_all(mode.wrap_unet)
    eps_unet = mode.wrap_guidance_unet(unet_cfg, guidance_scale, B)
    cscheduler.set_eps_unets([eps_unet])
    targs = {"strength": strength} if image is not None else {}
    cscheduler.set_timesteps(steps, config=cs.SchedulerConfig(eta=eta, karras_rho=karras_rho), prediction_type=prediction_type,
                             **targs)
    cscheduler.unet = mode.wrap_k_unet(cscheduler.unets[0])
    latents = mode.generateLatents()
    smod = __import__("inspect").getmodule(getattr(sampler_fn, "func", sampler_fn))
    keep = {k_: getattr(smod, k_) for k_ in ("torch", "trange", "tqdm") if hasattr(smod, k_)}
    try:
        with torch.no_grad():
            return cscheduler.loop(latents, lambda it: it)
    finally:
        for k_, v_ in keep.items():
            setattr(smod, k_, v_)


def pin_segment():
    """PINS the oracle's restatement of the host-side hot segment (oracle/sampling.py: txt2img_latents, image_mode_latents -
    generateLatents, the image modes, KDiffusionScheduler.set_timesteps / loop, the CFG wrapper stack) against the reference's
    own classes run here (scripts/_vendored.py:gyre_unified_pipeline loads gyre/pipeline/unified_pipeline.py and
    common_scheduler.py with absent third-party packages stood in for by empty classes).  Only the UNet and the VAE under the
    segment are the oracle's (diffusers is absent)."""
    up = _vendored.gyre_unified_pipeline()
    _, ksamp, _ = _vendored.k_diffusion()
    cfg = UNetConfig.tiny()
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    unet = OracleUNet(cfg, P)
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    unc = torch.randn(1, 77, cfg.cross_attention_dim, generator=torch.Generator().manual_seed(12)).expand(2, -1, -1)
    seeds = [420420420, 420420421]
    out = {}
    fns = {"euler_a": ksamp.sample_euler_ancestral, "euler": ksamp.sample_euler, "heun": ksamp.sample_heun,
           "dpm_2": ksamp.sample_dpm_2, "dpm_2_a": ksamp.sample_dpm_2_ancestral, "lms": ksamp.sample_lms,
           "dpmpp_2s_a": ksamp.sample_dpmpp_2s_ancestral, # samplers.py:58-60 registers gyre's DPM++ 2M with these two arguments bound
           "dpmpp_2m": __import__("functools").partial(_vendored.gyre_dpmpp_2m().sample_dpmpp_2m, warmup_lms=True, ddim_cutoff=0.1)}
    for (name, steps, hw, execution, ldt) in (("euler_a", 6, (128, 128), "parallel", torch.float32),
                                              ("euler_a", 5, (128, 128), "sequential", torch.float32),
                                              ("euler_a", 5, (128, 128), "parallel", torch.float16),
                                              ("euler", 5, (192, 128), "parallel", torch.float32),
                                              ("euler_a", 4, (64, 128), "parallel", torch.float32),
                                              ("heun", 4, (128, 128), "parallel", torch.float32),
                                              ("dpm_2", 4, (128, 128), "parallel", torch.float32),
                                              ("dpm_2_a", 4, (128, 128), "parallel", torch.float32),
                                              ("lms", 5, (128, 128), "parallel", torch.float32),
                                              ("dpmpp_2s_a", 4, (128, 128), "parallel", torch.float32),
                                              ("dpmpp_2m", 5, (128, 128), "parallel", torch.float32)):
        u_ = unet
        if ldt != torch.float32:
            class _Cast:                      # an fp32 UNet under fp16 latents (the CPU has no fp16 kernels for the blocks)
                config = unet.config

                def __call__(self, latents, t, **kw):
                    return OracleUNet._Out(unet(latents.float(), t, **{k: (v.float() if torch.is_tensor(v) else v)
                                                                       for k, v in kw.items()}).sample.to(latents.dtype))
            u_ = _Cast()
        ref = _reference_segment(up, u_, None, unc, emb, kind="txt2img", sampler_fn=fns[name], steps=steps, seeds=seeds,
                                 height=hw[0], width=hw[1], sample_size=16, cfg_execution=execution, latents_dtype=ldt)
        cfgu = (osamp.CFGParallel if execution == "parallel" else osamp.CFGSequential)(u_, unc, emb, 7.5)
        mine = osamp.txt2img_latents(cfgu, batch=2, in_channels=4, height=hw[0], width=hw[1], sample_size=16, seeds=seeds,
                                     steps=steps, sampler=name, latent_dtype=ldt)
        err = (ref.float() - mine.float()).abs().max().item() / ref.float().abs().max().item()
        key = f"txt2img/{name}/{steps}/{hw[0]}x{hw[1]}/{execution}/{str(ldt).split('.')[-1]}"
        print(f"  {key}: rel diff {err:.2e}")
        # (fp16 latents: the reference rounds every update to fp16, the oracle keeps the state in fp32 and reproduces the casts
        # that change the RESULT'S MEANING - sigma, noise draws - so this case agrees to fp16 rounding only)
        assert err < (2e-5 if ldt == torch.float32 else 4e-3), key
        out[key] = ref
    vcfg = VAEConfig.tiny()
    VP = synth_params(vae_param_shapes(vcfg), seed=4321)
    gi = torch.Generator().manual_seed(21)
    emb_i = torch.randn(2, 77, cfg.cross_attention_dim, generator=gi)
    unc_i = torch.randn(1, 77, cfg.cross_attention_dim, generator=gi).expand(2, -1, -1).contiguous()
    image = torch.rand(1, 3, 128, 128, generator=gi)
    mask = torch.zeros(1, 1, 128, 128)
    mask[:, :, 32:96, 40:104] = 1.0
    cfg9 = UNetConfig.tiny(in_channels=9)
    unet9 = OracleUNet(cfg9, synth_params(unet_param_shapes(cfg9), seed=1234))
    for kind, strength, u_ in (("img2img", 0.6, unet), ("runway", 0.75, unet9), ("runway", 1.0, unet9), ("inpaint", 0.8, unet)):
        vae = OracleVAE(vcfg, VP)
        kw = {} if kind == "img2img" else {"mask_image": mask}
        ref = _reference_segment(up, u_, vae, unc_i, emb_i, kind=kind, sampler_fn=ksamp.sample_euler_ancestral, steps=10,
                                 seeds=seeds, height=128, width=128, sample_size=16, image=image, strength=strength, **kw)
        mine = osamp.image_mode_latents(u_, OracleVAE(vcfg, VP), unc_i, emb_i, 7.5, image=image, seeds=seeds, steps=10,
                                        strength=strength, **kw)
        err = (ref.float() - mine.float()).abs().max().item() / ref.float().abs().max().item()
        key = f"{kind}/{strength}"
        print(f"  {key}: rel diff {err:.2e}")
        assert err < 2e-5, key
        out[key] = ref
    torch.save(out, os.path.join(GOLD, "segment.pt"))
    print(f"segment: {len(out)} runs of the reference's own mode / scheduler / CFG classes == oracle/sampling.py")


def pin_hint_classes():
    """PINS oracle/hints.py ControlnetHint / T2iHint (and with them the hint path of gyre_b200/hints.py, tested against the
    oracle on the GPU) against the reference's UnifiedPipelineHint_Controlnet / UnifiedPipelineHint_T2i run inside the
    reference's own wrapper stack, scheduler and Txt2imgMode (see pin_segment); the ControlNet / adapter / UNet under them are
    the oracle's."""
    from types import SimpleNamespace as SN
    from oracle import controlnet as ocn
    from oracle import hints as oh
    from oracle import t2i_adapter as oad
    up = _vendored.gyre_unified_pipeline()
    _, ksamp, _ = _vendored.k_diffusion()
    cfg = UNetConfig.tiny()
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    Pcn = synth_params(ocn.controlnet_param_shapes(cfg), seed=77)
    akw = dict(channels=list(cfg.block_out_channels), nums_rb=2, cin=192, ksize=1, sk=True, use_conv=False)
    Pad = synth_params(oad.adapter_param_shapes(**akw), seed=91)
    unet = OracleUNet(cfg, P)
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    unc = torch.randn(1, 77, cfg.cross_attention_dim, generator=g).expand(2, -1, -1).contiguous()
    img = torch.rand(1, 3, 128, 128, generator=g).half().float()

    class CN:
        config = {}

        def __call__(self, cnlatents, t, encoder_hidden_states, controlnet_cond):
            with torch.no_grad():
                down, mid = ocn.controlnet_forward(Pcn, cfg, cnlatents, t, encoder_hidden_states, controlnet_cond)
            return SN(down_block_res_samples=down, mid_block_res_sample=mid)

    class _Cfg(dict):
        __getattr__ = dict.__getitem__

    class AD:
        config = _Cfg(cin=192)

        def __call__(self, x):
            with torch.no_grad():
                return oad.adapter_forward(Pad, x, **{k: v for k, v in akw.items() if k != "cin"})

        def parameters(self):
            return iter([torch.zeros(1)])

    # hint masks: an RGBA hint (mask = alpha) at 256 x 256 - the 1 / 64 antialiased downscale for the deepest residual needs
    # 192 pixels of reflect padding, more than a 128-pixel mask has (the reference fails there) - and a ControlNet under the
    # 9-channel inpaint UNet, whose mask comes from the UNet input
    big = torch.rand(1, 3, 256, 256, generator=g).half().float()
    alpha = torch.zeros(1, 1, 256, 256)
    alpha[:, :, 64:224, 32:160] = 1.0
    rgba = torch.cat([big, alpha], dim=1)
    cfg9 = UNetConfig.tiny(in_channels=9)
    unet9 = OracleUNet(cfg9, synth_params(unet_param_shapes(cfg9), seed=1234))
    vcfg = VAEConfig.tiny()
    VP = synth_params(vae_param_shapes(vcfg), seed=4321)
    out = {"rgba": rgba.half()}
    for name, kind in (("masked controlnet + t2i", "txt2img"), ("controlnet under runway inpaint", "runway")):
        if kind == "txt2img":
            rh = [up.UnifiedPipelineHint_Controlnet(CN(), rgba.clone(), None, 0.9, True, False, batch_total=2),
                  up.UnifiedPipelineHint_T2i(AD(), rgba.expand(2, -1, -1, -1).clone(), None, None, None, None, 0.8, False, False, None)]
            ohints = [oh.ControlnetHint(CN(), rgba, weight=0.9, soft_injection=True, cfg_only=False),
                      oh.T2iHint(AD(), rgba.expand(2, -1, -1, -1), weight=0.8, soft_injection=False, cfg_only=False)]
            kw, u_ = {}, unet
        else:
            rh = [up.UnifiedPipelineHint_Controlnet(CN(), rgba.clone(), None, 1.0, False, False, batch_total=2)]
            ohints = [oh.ControlnetHint(CN(), rgba, weight=1.0, soft_injection=False, cfg_only=False)]
            img_in = torch.rand(1, 3, 256, 256, generator=g).half().float()
            msk_in = torch.zeros(1, 1, 256, 256)
            msk_in[:, :, 96:192, 64:200] = 1.0
            kw, u_ = dict(image=img_in, mask_image=msk_in, strength=0.8), unet9
            out["runway_image"], out["runway_mask"] = img_in.half(), msk_in.half()
        for h in rh:
            h.to(torch.device("cpu"), torch.float32)
        ref = _reference_segment(up, u_, OracleVAE(vcfg, VP), unc, emb, kind=kind, sampler_fn=ksamp.sample_euler_ancestral, steps=4,
                                 seeds=[420420420, 420420421], height=256, width=256, sample_size=16, hints=rh, **kw)
        if kind == "txt2img":
            eps = oh.guided_eps_unet(u_, unc, emb, 7.5, ohints)
            with torch.no_grad():
                mine = osamp.txt2img_latents(eps, batch=2, in_channels=4, height=256, width=256, sample_size=16,
                                             seeds=[420420420, 420420421], steps=4, sampler="euler_a")
        else:
            with torch.no_grad():
                mine = osamp.image_mode_latents(u_, OracleVAE(vcfg, VP), unc, emb, 7.5, seeds=[420420420, 420420421], steps=4,
                                                hints=ohints, **kw)
        err = (ref - mine).abs().max().item() / ref.abs().max().item()
        print(f"  hints/{name}: rel diff {err:.2e}")
        assert err < 2e-5, name
        out[name] = ref
    # a style adapter (`style_call`, unified_pipeline.py:941-975: images.rescale "cover", CLIP normalisation, the CLIP vision
    # tower's hidden state of the chosen layer, StyleAdapter) next to a standard adapter (alone, the reference's UNetWithT2I
    # reads an attribute it only sets for standard states, core.py:201, 214)
    from oracle import safety as osf
    sg = torch.load(os.path.join(GOLD, "safety.pt"))["models"]["tiny"]
    vis = sg["vision_config"]
    Pv = {k[len("vision_model."):]: v.float() for k, v in sg["state_dict"].items() if k.startswith("vision_model.vision_model.")}
    skw = dict(width=vis["hidden_size"], context_dim=cfg.cross_attention_dim, num_head=4, n_layes=2, num_token=4)
    from gyre_b200.clip_vision import style_adapter_param_shapes
    Pst = synth_params(style_adapter_param_shapes(**skw), seed=47)

    def vision(image, output_hidden_states=False, return_dict=True):
        with torch.no_grad():
            _, _, hs = osf.clip_vision_forward(Pv, image, num_layers=vis["num_hidden_layers"], num_heads=vis["num_attention_heads"],
                                               patch_size=vis["patch_size"], hidden_act=vis["hidden_act"], return_hidden_states=True)
        return SN(last_hidden_state=hs[-1], hidden_states=tuple(hs))

    class ST(up.t2i_adapter.T2iAdapter_style):
        config = _Cfg()
        _coadapter_type = False

        def __call__(self, x):
            with torch.no_grad():
                return oad.style_adapter_forward(Pst, x, num_head=skw["num_head"], num_token=skw["num_token"])
    up.modeling_utils.get_parameter_dtype = lambda m_: torch.float32
    fe = SN(image_mean=list(osf.CLIP_MEAN), image_std=list(osf.CLIP_STD), size={"shortest_edge": vis["image_size"]})
    style_img = torch.rand(1, 3, 96, 128, generator=g).half().float()
    out["style_image"] = style_img.half()
    for name, layer in (("style final + t2i", None), ("style penultimate + t2i", "penultimate")):
        rh = [up.UnifiedPipelineHint_T2i(AD(), img.expand(2, -1, -1, -1).clone(), None, None, None, None, 1.0, False, False, None),
              up.UnifiedPipelineHint_T2i(ST(), style_img.expand(2, -1, -1, -1).clone(), SN(vision_model=vision), fe, None, None, 0.7,
                                         False, True, layer)]
        for h in rh:
            h.to(torch.device("cpu"), torch.float32)
        ref = _reference_segment(up, unet, None, unc, emb, kind="txt2img", sampler_fn=ksamp.sample_euler_ancestral, steps=4,
                                 seeds=[420420420, 420420421], height=128, width=128, sample_size=16, hints=rh)
        style = dict(vision=vision, mean=osf.CLIP_MEAN, std=osf.CLIP_STD, size=vis["image_size"], clip_layer=layer)
        ohints = [oh.T2iHint(AD(), img.expand(2, -1, -1, -1)),
                  oh.T2iHint(ST(), style_img.expand(2, -1, -1, -1), weight=0.7, cfg_only=True, style=style)]
        with torch.no_grad():
            mine = osamp.txt2img_latents(oh.guided_eps_unet(unet, unc, emb, 7.5, ohints), batch=2, in_channels=4, height=128,
                                         width=128, sample_size=16, seeds=[420420420, 420420421], steps=4, sampler="euler_a")
        err = (ref - mine).abs().max().item() / ref.abs().max().item()
        print(f"  hints/{name}: rel diff {err:.2e}")
        assert err < 2e-5, name
        out[name] = ref
    seeds, steps = [420420420, 420420421], 5
    cases = [("controlnet", dict(weight=1.0, soft_injection=False, cfg_only=False), None, "parallel"),
             ("controlnet soft 0.7", dict(weight=0.7, soft_injection=True, cfg_only=False), None, "parallel"),
             ("controlnet cfg_only", dict(weight=0.8, soft_injection=True, cfg_only=True), None, "parallel"),
             ("controlnet cfg_only sequential", dict(weight=0.8, soft_injection=False, cfg_only=True), None, "sequential"),
             ("t2i", None, dict(weight=1.0, soft_injection=False, cfg_only=False), "parallel"),
             ("t2i soft cfg_only + controlnet", dict(weight=0.5, soft_injection=False, cfg_only=False),
              dict(weight=0.9, soft_injection=True, cfg_only=True), "parallel"),
             ("t2i cfg_only sequential", None, dict(weight=1.0, soft_injection=False, cfg_only=True), "sequential")]
    for name, cn_kw, ad_kw, execution in cases:
        rh, ohints = [], []
        if cn_kw is not None:
            rh.append(up.UnifiedPipelineHint_Controlnet(CN(), img, None, cn_kw["weight"], cn_kw["soft_injection"],
                                                        cn_kw["cfg_only"], batch_total=2))
            ohints.append(oh.ControlnetHint(CN(), img, **cn_kw))
        if ad_kw is not None:
            # (the reference adds adapter states to the hidden states as they are: hint batch = sample batch)
            rh.append(up.UnifiedPipelineHint_T2i(AD(), img.expand(2, -1, -1, -1), None, None, None, None, ad_kw["weight"],
                                                 ad_kw["soft_injection"], ad_kw["cfg_only"], None))
            ohints.append(oh.T2iHint(AD(), img.expand(2, -1, -1, -1), **ad_kw))
        for h in rh:
            h.to(torch.device("cpu"), torch.float32)
        ref = _reference_segment(up, unet, None, unc, emb, kind="txt2img", sampler_fn=ksamp.sample_euler_ancestral, steps=steps,
                                 seeds=seeds, height=128, width=128, sample_size=16, cfg_execution=execution, hints=rh)
        eps = oh.guided_eps_unet(unet, unc, emb, 7.5, ohints, parallel=execution == "parallel")
        with torch.no_grad():
            mine = osamp.txt2img_latents(eps, batch=2, in_channels=4, height=128, width=128, sample_size=16, seeds=seeds,
                                         steps=steps, sampler="euler_a")
        err = (ref - mine).abs().max().item() / ref.abs().max().item()
        print(f"  hints/{name}: rel diff {err:.2e}")
        assert err < 2e-5, name
        out[name] = ref
    torch.save(out, os.path.join(GOLD, "hint_classes.pt"))
    print(f"hint classes: {len([k for k in out if not k.startswith('r') and k != 'style_image'])} runs of UnifiedPipelineHint_Controlnet / _T2i inside the reference's stack == oracle/hints.py")


def _reference_call(up, *, unet, vae, unc, emb, sampler_fn, seeds, inpaint_unet=None, depth_unet=None, options=None,
                    sample_size=16, **call_kwargs):
    """UnifiedPipeline.__call__ ITSELF (unified_pipeline.py:1722-2531) on an instance assembled without `__init__` (which
    registers diffusers modules): mode-tree construction with its hires-fix / graft decisions, the per-leaf wrapper stacks,
    scheduler, loop, split.  Stand-ins: the text-embedding calculator returns the given embeddings (LPW is pinned on its own),
    `vae_decode` captures what it is handed (= final latents / 0.18215) and returns a blank image."""
    from types import SimpleNamespace as SN
    up.DiffusionPipeline.device = torch.device("cpu")
    pipe = object.__new__(up.UnifiedPipeline)
    pipe.unet, pipe.vae, pipe.inpaint_unet, pipe.depth_unet = unet, vae, inpaint_unet, depth_unet
    pipe.text_encoder = pipe.inpaint_text_encoder = pipe.depth_text_encoder = None
    pipe.tokenizer = pipe.clip_tokenizer = SN(model_max_length=77)
    pipe.clip_model = pipe.safety_checker = pipe.feature_extractor = pipe.hintset_manager = None
    pipe.scheduler = sampler_fn
    pipe.vae_scale_factor, pipe.vae_dtype = 8, torch.float32
    pipe._grafted_depth = pipe._grafted_inpaint = False
    pipe._hires_fix, pipe._hires_threshold_fraction, pipe._hires_oos_fraction, pipe._hires_image_oos_fraction = True, 0.0333, 0.6, 1.0
    pipe._structured_diffusion, pipe._text_embedding_layer = False, "final"
    pipe.clip_default_config = SN(guidance_scale=0, guidance_base="guided", gradient_length=15, gradient_threshold=0.01,
                                  gradient_maxloss=1.0, vae_cutouts=2, approx_cutouts=2, no_cutouts=False)
    for k_, v_ in (options or {}).items():
        setattr(pipe, k_, v_)
    pipe.set_tiling_mode = lambda tiling: None
    pipe.progress_bar = lambda it: it
    pipe.get_unet_sample_size = lambda u: sample_size          # (the reference forces >= 64; the tiny UNets are 16)
    captured = {}

    def vae_decode(latents):
        captured["z"] = latents.clone()
        return torch.zeros(latents.shape[0], 3, latents.shape[2] * 8, latents.shape[3] * 8)
    pipe.vae_decode = vae_decode

    class Emb:                                                 # LPWTextEmbedding's surface as __call__ uses it (:2269-2304)
        def __init__(self, **kw):
            pass

        def get_embeddings(self, prompt, uncond_prompt=None):
            return emb, (unc if uncond_prompt is not None else None)

        def repeat(self, x, n):
            return x
    saved = up.LPWTextEmbedding
    up.LPWTextEmbedding = Emb
    # __call__ patches `torch` / `trange` / `tqdm` inside the sampler's module (gyre/patching.py): put them back afterwards
    smod = __import__("inspect").getmodule(getattr(sampler_fn, "func", sampler_fn))
    keep = {k_: getattr(smod, k_) for k_ in ("torch", "trange", "tqdm") if hasattr(smod, k_)}
    try:
        pipe(prompt=["a"] * len(seeds), generator=[torch.Generator("cpu").manual_seed(s_) for s_ in seeds],
             scheduler=sampler_fn, output_type="tensor", return_dict=False, run_safety_checker=False, **call_kwargs)
    finally:
        up.LPWTextEmbedding = saved
        for k_, v_ in keep.items():
            setattr(smod, k_, v_)
    return captured["z"]


def pin_call():
    """PINS the oracle's compositions of a whole request (oracle/hires.py hires_txt2img_latents / hires_image_mode_latents /
    grafted_inpaint_latents, oracle/sampling.py txt2img_latents / image_mode_latents) against UnifiedPipeline.__call__ itself,
    run here from /root/reference (see _reference_call) over the oracle UNets / VAE."""
    from oracle import hires as ohires
    up = _vendored.gyre_unified_pipeline()
    _, ksamp, _ = _vendored.k_diffusion()
    cfg = UNetConfig.tiny()
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    unet = OracleUNet(cfg, P)
    cfg9 = UNetConfig.tiny(in_channels=9)
    unet9 = OracleUNet(cfg9, synth_params(unet_param_shapes(cfg9), seed=1234))
    vcfg = VAEConfig.tiny()
    VP = synth_params(vae_param_shapes(vcfg), seed=4321)
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    unc = torch.randn(1, 77, cfg.cross_attention_dim, generator=torch.Generator().manual_seed(12)).expand(2, -1, -1).contiguous()
    seeds = [420420420, 420420421]
    out = {}

    def check(key, ref_z, mine):
        ref = 0.18215 * ref_z
        err = (ref - mine).abs().max().item() / mine.abs().max().item()
        print(f"  call/{key}: rel diff {err:.2e}")
        assert err < 2e-6, key
        out[key] = ref

    # txt2img at the native size (no hires fix: below the threshold)
    z = _reference_call(up, unet=unet, vae=OracleVAE(vcfg, VP), unc=unc, emb=emb, sampler_fn=ksamp.sample_euler_ancestral,
                        seeds=seeds, height=128, width=128, num_inference_steps=6)
    mine = osamp.txt2img_latents(osamp.CFGParallel(unet, unc, emb, 7.5), batch=2, in_channels=4, height=128, width=128,
                                 sample_size=16, seeds=seeds, steps=6, sampler="euler_a")
    check("txt2img 128", z, mine)

    # above the native size: the hires fix engages by default (natural-size twin, HiresUnetWrapper)
    for (H, W) in ((192, 192), (256, 192)):
        z = _reference_call(up, unet=unet, vae=OracleVAE(vcfg, VP), unc=unc, emb=emb, sampler_fn=ksamp.sample_euler_ancestral,
                            seeds=seeds, height=H, width=W, num_inference_steps=5)
        mine = ohires.hires_txt2img_latents(osamp.CFGParallel(unet, unc, emb, 7.5), batch=2, height=H, width=W, sample_size=16,
                                            seeds=seeds, steps=5, oos_fraction=0.6)
        check(f"hires txt2img {H}x{W}", z, mine)
    # ... and hires_fix=False gives the plain single-leaf run
    z = _reference_call(up, unet=unet, vae=OracleVAE(vcfg, VP), unc=unc, emb=emb, sampler_fn=ksamp.sample_euler_ancestral,
                        seeds=seeds, height=192, width=192, num_inference_steps=5, hires_fix=False)
    mine = osamp.txt2img_latents(osamp.CFGParallel(unet, unc, emb, 7.5), batch=2, in_channels=4, height=192, width=192,
                                 sample_size=16, seeds=seeds, steps=5, sampler="euler_a")
    check("txt2img 192 hires off", z, mine)

    gi = torch.Generator().manual_seed(21)
    image = torch.rand(1, 3, 128, 128, generator=gi)
    mask = torch.zeros(1, 1, 128, 128)
    mask[:, :, 32:96, 40:104] = 1.0
    big_image = torch.rand(1, 3, 192, 192, generator=gi)
    big_mask = torch.zeros(1, 1, 192, 192)
    big_mask[:, :, 48:144, 60:156] = 1.0
    # image modes through __call__ (mode selection :2055-2066): img2img, Runway inpaint (9-channel UNet), legacy inpaint
    for key, u_, kw in (("img2img", unet, dict(image=image, strength=0.6)),
                        ("runway inpaint", unet9, dict(image=image, mask_image=mask, strength=0.75)),
                        ("legacy inpaint", unet, dict(image=image, mask_image=mask, strength=0.8))):
        z = _reference_call(up, unet=u_, vae=OracleVAE(vcfg, VP), unc=unc, emb=emb, sampler_fn=ksamp.sample_euler_ancestral,
                            seeds=seeds, inpaint_unet=u_ if u_ is unet9 else None, height=128, width=128, num_inference_steps=8, **kw)
        mine = osamp.image_mode_latents(u_, OracleVAE(vcfg, VP), unc, emb, 7.5, seeds=seeds, steps=8,
                                        **{("mask_image" if k == "mask_image" else k): v for k, v in kw.items()})
        check(key, z, mine)
    # grafted inpaint (:2069-2098): inpaint UNet early, main UNet with the legacy blend late
    z = _reference_call(up, unet=unet, vae=OracleVAE(vcfg, VP), unc=unc, emb=emb, sampler_fn=ksamp.sample_euler_ancestral,
                        seeds=seeds, inpaint_unet=unet9, options={"_grafted_inpaint": True}, height=128, width=128,
                        num_inference_steps=8, image=image, mask_image=mask, strength=0.75)
    mine = ohires.grafted_inpaint_latents(unet9, unet, OracleVAE(vcfg, VP), unc, emb, 7.5, image=image, mask_image=mask,
                                          seeds=seeds, steps=8, strength=0.75)
    check("grafted inpaint", z, mine)
    # hires fix over the image modes (oos fraction 1.0 when an image is given, :1840-1843)
    for key, u_, kw in (("hires img2img", unet, dict(image=big_image, strength=0.6)),
                        ("hires runway inpaint", unet9, dict(image=big_image, mask_image=big_mask, strength=0.75))):
        z = _reference_call(up, unet=u_, vae=OracleVAE(vcfg, VP), unc=unc, emb=emb, sampler_fn=ksamp.sample_euler_ancestral,
                            seeds=seeds, inpaint_unet=u_ if u_ is unet9 else None, height=192, width=192, num_inference_steps=6, **kw)
        mine = ohires.hires_image_mode_latents(u_, OracleVAE(vcfg, VP), unc, emb, 7.5, seeds=seeds, steps=6, sample_size=16,
                                               oos_fraction=1.0, **kw)
        check(key, z, mine)
    # a depth hint: the request goes to the 5-channel depth UNet (:2004-2013), optionally grafted onto the main UNet
    gim = sys.modules["gyre.images"]
    pt = sys.modules["gyre.pipeline.prompt_types"]
    cfg5 = UNetConfig.tiny(in_channels=5)
    unet5 = OracleUNet(cfg5, synth_params(unet_param_shapes(cfg5), seed=321))
    # (UnetWithExtraChannels takes the channels as they are: the hint batch has to equal the sample batch, core.py:22-25)
    depth_img = torch.rand(1, 1, 128, 128, generator=torch.Generator().manual_seed(31)).expand(2, -1, -1, -1).contiguous()
    depth_map = 2.0 * gim.resize(gim.normalise_tensor(depth_img, 1), (1 / 8, 1 / 8), sharpness=2) - 1.0     # the reference's lines
    out["depth_map"] = depth_map
    out["depth_image"] = depth_img
    blend = {"start": 0.15, "end": 0.75, "easing": "sine"}
    for key, opts, gb in (("depth", {}, None), ("grafted depth", {"_grafted_depth": blend}, blend)):
        z = _reference_call(up, unet=unet, vae=OracleVAE(vcfg, VP), unc=unc, emb=emb, sampler_fn=ksamp.sample_euler_ancestral,
                            seeds=seeds, depth_unet=unet5, options=opts, height=128, width=128, num_inference_steps=7,
                            hint_images=[pt.HintImage(image=depth_img, hint_type="depth")])
        mine = ohires.depth_txt2img_latents(unet5, unet, unc, emb, 7.5, depth_map=depth_map, seeds=seeds, steps=7,
                                            sample_size=16, height=128, width=128, graft_blend=gb)
        check(key, z, mine)
    # a diffusers-protocol scheduler (DiffusersScheduler wrapper, common_scheduler.py:179-331) around the in-tree DDIM copy
    ddim = _vendored.gyre_ddim().DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                              clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    z = _reference_call(up, unet=unet, vae=OracleVAE(vcfg, VP), unc=unc, emb=emb, sampler_fn=ddim, seeds=seeds, height=128,
                        width=128, num_inference_steps=8)
    mine = osamp.txt2img_latents(osamp.CFGParallel(unet, unc, emb, 7.5), batch=2, in_channels=4, height=128, width=128,
                                 sample_size=16, seeds=seeds, steps=8, sampler="ddim")
    check("ddim", z, mine)
    torch.save(out, os.path.join(GOLD, "call.pt"))
    print(f"call: {len(out) - 2} runs of UnifiedPipeline.__call__ == the oracle's compositions")


def pin_resize():
    """PINS oracle/hires.py:images_resize (and through tests/test_images_gpu.py the native gyre_b200.images.resize) against the
    reference's gyre/images.py:resize over the vendored ResizeRight."""
    from oracle import hires as ohires
    _vendored.gyre_unified_pipeline()
    gim = sys.modules["gyre.images"]
    g = torch.Generator().manual_seed(0)
    out = []
    for shape, f, sh in [((1, 1, 128, 128), (1 / 8, 1 / 8), 1), ((1, 1, 128, 128), (1 / 8, 1 / 8), 2), ((2, 1, 16, 16), (8, 8), 1),
                         ((1, 1, 96, 128), (0.25, 0.125), 1), ((1, 3, 40, 24), (2.0, 1.0), 1), ((1, 1, 256, 256), (1 / 32, 1 / 32), 1),
                         ((1, 1, 64, 64), (0.5, 0.5), 2), ((1, 1, 128, 128), (1 / 16, 1 / 16), 1), ((2, 1, 16, 24), (8, 8), 2)]:
        x = torch.rand(shape, generator=g)
        if shape[-1] == 16:
            x = (x > 0.5).float()                       # a hard mask: the lanczos overshoot meets the clamp
        ref = gim.resize(x, f, sharpness=sh)
        mine = ohires.images_resize(x, f, sharpness=sh)
        err = (ref - mine).abs().max().item()
        assert ref.shape == mine.shape and err <= 5e-7, (shape, f, sh, err)
        out.append({"x": x, "factors": f, "sharpness": sh, "out": ref})
    torch.save(out, os.path.join(GOLD, "resize.pt"))
    print(f"resize: {len(out)} cases of gyre.images.resize == oracle (<= 5e-7: torch.sum's summation order)")


def oracle_fixtures(full: bool):
    """Oracle self-fixtures (unpinned at the diffusers boundary)."""
    out = {}
    cfg = UNetConfig.tiny()
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    g = torch.Generator("cpu").manual_seed(5)
    x = torch.randn(2, 4, 16, 16, generator=g)
    ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    t = torch.tensor([981, 21])
    with torch.no_grad():
        out["unet_tiny"] = {"x": x, "ctx": ctx, "t": t, "eps": unet_forward(P, cfg, x, t, ctx)}
        out["unet_tiny_tome"] = {"r": 48, "eps": unet_forward(P, cfg, x, t, ctx, tome_r=48)}
        vcfg = VAEConfig.tiny()
        VP = synth_params(vae_param_shapes(vcfg), seed=4321)
        z = torch.randn(1, 4, 8, 8, generator=g)
        out["vae_tiny"] = {"z": z, "img": vae_decode(VP, vcfg, z)}
        img = torch.rand(1, 3, 32, 32, generator=g) * 2 - 1
        out["vae_tiny_enc"] = {"img": img, "moments": vae_encode_moments(VP, vcfg, img)}
        unet = OracleUNet(cfg, P)
        emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=torch.Generator().manual_seed(11))
        unc = torch.randn(1, 77, cfg.cross_attention_dim, generator=torch.Generator().manual_seed(12)).expand(2, -1, -1)
        cfgu = osamp.CFGParallel(unet, unc, emb, 7.5)
        for sampler, steps in (("ddim", 10), ("euler_a", 12), ("euler", 8), ("dpmpp_2m", 8), ("heun", 7), ("dpm_2", 7),
                               ("dpm_2_a", 7), ("lms", 9), ("dpmpp_2s_a", 7), ("dpmpp_sde", 7), ("dpm_fast", 10), ("plms", 9),
                               ("dpmsolverpp_1", 9), ("dpmsolverpp_2", 9), ("dpmsolverpp_3", 11), ("dpmsolverpp_3b", 20)):
            lat = osamp.txt2img_latents(cfgu, batch=2, in_channels=4, height=128, width=128, sample_size=16,
                                        seeds=[420420420, 420420421], steps=steps, sampler=sampler.rstrip("b"))
            out[f"pipe_tiny/{sampler}"] = {"steps": steps, "latents": lat}
    torch.save(out, os.path.join(GOLD, "oracle_tiny.pt"))
    print("oracle tiny fixtures:", list(out))
    if full:
        import time
        cfg = UNetConfig.sd15()
        P = synth_params(unet_param_shapes(cfg), seed=1234)
        unet = OracleUNet(cfg, P)
        emb = torch.randn(1, 77, 768, generator=torch.Generator().manual_seed(11))
        unc = torch.randn(1, 77, 768, generator=torch.Generator().manual_seed(12))
        cfgu = osamp.CFGParallel(unet, unc, emb, 7.5)
        t0 = time.time()
        with torch.no_grad():
            lat = osamp.txt2img_latents(cfgu, batch=1, in_channels=4, height=512, width=512, sample_size=64,
                                        seeds=[420420420], steps=10, sampler="ddim")
        print("C1 full-size oracle run: %.1fs" % (time.time() - t0))
        torch.save({"latents": lat, "steps": 10, "sampler": "ddim", "seed": 420420420, "weights_seed": 1234},
                   os.path.join(GOLD, "c1_sd15_ddim10.pt"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated subset: samplers,ddim,tome,clip,wrappers,hires,lpw,images,t2i_adapter,safety,controlnet,hints,attention,segment,hint_classes,call,resize,oracle")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    parts = {"samplers": pin_samplers, "ddim": pin_ddim, "tome": pin_tome, "clip": pin_clip, "wrappers": pin_wrappers,
             "hires": pin_hires, "lpw": pin_lpw, "images": pin_images, "t2i_adapter": pin_t2i_adapter, "safety": pin_safety, "controlnet": pin_controlnet, "hints": pin_hints, "attention": pin_attention, "segment": pin_segment, "hint_classes": pin_hint_classes, "call": pin_call, "resize": pin_resize, "oracle": lambda: oracle_fixtures(a.full)}
    for name, fn in parts.items():
        if not a.only or name in a.only.split(","):
            fn()
