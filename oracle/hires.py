"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's hires-fix and graft scheduler-UNet wrappers.

Follows gyre/pipeline/unet/hires_fix.py:22-235 (`pad_like`, `scale_into`, `down_scale_factor`, `up_scale_factor`,
`scale_strategy`, `HiresUnetWrapper`), gyre/pipeline/unet/graft.py:16-56 (`GraftUnets`), gyre/pipeline/easing.py:22-49
(`Easing`) and the vendored ResizeRight (gyre/src/ResizeRight/resize_right.py:30-127, 203-252 - separable resampling
with a 4-tap lanczos2 window, interp_methods.py:47-51) for the one call shape gyre uses: a scalar scale factor on the
last two dims, `antialiasing=False`, `pad_mode="replicate"`, `by_convs=False`.

The easing curves themselves live in a third-party dependency that is absent here - `easing-functions ~= 1.0.4`
(pyproject.toml:24): `EasingBase.ease(alpha)` = end * f(t) + start * (1 - f(t)), t = alpha / duration, with the
published Penner in-out curves restated below.  Everything else is PINNED: scripts/make_golden.py runs the reference's
own hires_fix.py / graft.py / easing.py / resize_right.py (with these curves standing in for the absent package) and
asserts this restatement reproduces them bit for bit (tests/golden/hires.pt); the request-level compositions
(hires_txt2img_latents, hires_image_mode_latents, grafted_inpaint_latents) are pinned against UnifiedPipeline.__call__
itself - the reference's mode tree deciding about the hires fix and the graft (scripts/make_golden.py:pin_call,
tests/golden/call.pt).

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module.
"""
from __future__ import annotations

import math

import torch

Hi, Wi = -2, -1


# --------------------------------------------------------------------------------------------- easing-functions 1.0.4
class EasingBase:
    limit = (0, 1)

    def __init__(self, start=0, end=1, duration=1):
        self.start, self.end, self.duration = start, end, duration

    def func(self, t):
        raise NotImplementedError

    def ease(self, alpha):
        t = self.limit[0] * (1 - alpha) + self.limit[1] * alpha
        t /= self.duration
        a = self.func(t)
        return self.end * a + self.start * (1 - a)

    def __call__(self, alpha):
        return self.ease(alpha)


class LinearInOut(EasingBase):
    def func(self, t):
        return t


class QuadEaseInOut(EasingBase):
    def func(self, t):
        if t < 0.5:
            return 2 * t * t
        return (-2 * t * t) + (4 * t) - 1


class CubicEaseInOut(EasingBase):
    def func(self, t):
        if t < 0.5:
            return 4 * t * t * t
        p = 2 * t - 2
        return 0.5 * p * p * p + 1


class QuarticEaseInOut(EasingBase):
    def func(self, t):
        if t < 0.5:
            return 8 * t * t * t * t
        p = t - 1
        return -8 * p * p * p * p + 1


class QuinticEaseInOut(EasingBase):
    def func(self, t):
        if t < 0.5:
            return 16 * t * t * t * t * t
        p = (2 * t) - 2
        return 0.5 * p * p * p * p * p + 1


class SineEaseInOut(EasingBase):
    def func(self, t):
        return 0.5 * (1 - math.cos(t * math.pi))


class CircularEaseInOut(EasingBase):
    def func(self, t):
        if t < 0.5:
            return 0.5 * (1 - math.sqrt(1 - 4 * (t * t)))
        return 0.5 * (math.sqrt(-((2 * t) - 3) * ((2 * t) - 1)) + 1)


class ExponentialEaseInOut(EasingBase):
    def func(self, t):
        if t == 0 or t == 1:
            return t
        if t < 0.5:
            return 0.5 * math.pow(2, (20 * t) - 10)
        return -0.5 * math.pow(2, (-20 * t) + 10) + 1


EASINGS = {"linear": LinearInOut, "quad": QuadEaseInOut, "cubic": CubicEaseInOut, "quartic": QuarticEaseInOut,
           "quintic": QuinticEaseInOut, "sine": SineEaseInOut, "circular": CircularEaseInOut, "expo": ExponentialEaseInOut}


class Easing:
    """gyre/pipeline/easing.py:22-49."""

    def __init__(self, floor, start, end, easing):
        self.floor, self.start, self.end = floor, start, end
        if isinstance(easing, str):
            easing = EASINGS[easing]
        self.easing = easing(end=1 - floor, duration=(end - start))

    def interp(self, u):
        if u < self.start:
            return self.floor
        if u > self.end:
            return 1
        return self.floor + self.easing(u - self.start)


# --------------------------------------------------------------------------------------------- ResizeRight, lanczos2
def lanczos2(x):
    """interp_methods.py:47-51 (fp32 in, fp32 out)."""
    eps = torch.finfo(torch.float32).eps
    return ((torch.sin(math.pi * x) * torch.sin(math.pi * x / 2) + eps) / ((math.pi ** 2 * x ** 2 / 2) + eps)) * (abs(x) < 2).to(x.dtype)


def resample_taps(in_sz: int, scale: float):
    """The 1-D plan of resize_right.py:70-118 for one dim: out_sz = ceil(scale * in_sz); per output position the 4
    source indices (replicate padding == clamped indices) and normalised lanczos2 weights (fp32)."""
    eps = torch.finfo(torch.float32).eps
    out_sz = math.ceil(scale * in_sz)
    out_coords = torch.arange(out_sz)
    grid = out_coords / float(scale) + (in_sz - 1) / 2 - (out_sz - 1) / (2 * float(scale))      # get_projected_grid :128-141
    left = (grid - 4 / 2 - eps).ceil().long()                                                     # get_field_of_view :144-154
    fov = left[:, None] + torch.arange(math.ceil(4 - eps))
    # calc_pad_sz :157-172: the (generalised, possibly negative) left pad shifts BOTH the field of view and the fp32
    # grid before the weights are formed - the shift is part of the weights' rounding
    pad0 = -fov[0, 0].item()
    w = lanczos2((grid + pad0)[:, None] - (fov + pad0))                                           # get_weights :203-213
    s = w.sum(1, keepdim=True)
    s[s == 0] = 1
    w = w / s
    return fov.clamp(0, in_sz - 1), w, out_sz


def resize_lanczos2(x, scale: float):
    """resize_right.resize(x, scale_factors=scale, interp_method=lanczos2, pad_mode="replicate", antialiasing=False)
    followed by gyre's cast back to the input dtype (gyre/resize_right.py:41-42).  Dims are processed in ascending
    scale order (:55-59) - equal scales keep H before W - and a scale of exactly 1 is skipped."""
    if float(scale) == 1.0:
        return x
    out = x
    for dim in (-2, -1):
        idx, w, _ = resample_taps(out.shape[dim], scale)
        t = out.transpose(dim, 0)                       # apply_weights :216-250
        nb = t[idx]                                     # [out, 4, ...]
        ww = w.reshape(*w.shape, *([1] * (t.ndim - 1)))
        out = (nb * ww).sum(1).transpose(0, dim)
    return out.to(x.dtype)


def scale_into(latents, scale, target=None, target_shape=None):
    """hires_fix.py:45-92 (mode "lanczos")."""
    latents = resize_lanczos2(latents, scale)
    if (target is None) == (target_shape is None):
        raise ValueError("exactly one of target or target_shape")
    if target_shape is None:
        target_shape = target.shape
    offh = (target_shape[Hi] - latents.shape[Hi]) // 2
    offw = (target_shape[Wi] - latents.shape[Wi]) // 2
    if offh < 0:
        latents = latents[:, :, -offh:-offh + target_shape[Hi], :]
        offh = 0
    if offw < 0:
        latents = latents[:, :, :, -offw:-offw + target_shape[Wi]]
        offw = 0
    if target is not None:
        target[:, :, offh:offh + latents.shape[Hi], offw:offw + latents.shape[Wi]] = latents
        return target
    pad_w = [offw, (target_shape[Wi] - latents.shape[Wi]) - offw]
    pad_h = [offh, (target_shape[Hi] - latents.shape[Hi]) - offh]
    return torch.nn.functional.pad(latents, pad_w + pad_h, mode="replicate")


def down_scale_factor(latents_shape, target_shape, oos_fraction):
    """hires_fix.py:95-99."""
    scales = target_shape[Hi] / latents_shape[Hi], target_shape[Wi] / latents_shape[Wi]
    scale_min, scale_max = min(*scales), max(*scales)
    return scale_min * oos_fraction + scale_max * (1 - oos_fraction)


def up_scale_factor(latents_shape, target_shape, oos_fraction):
    return 1 / down_scale_factor(target_shape, latents_shape, oos_fraction)


def batched_rand(shape, generators, device, dtype):
    """gyre/pipeline/randtools.py:11-36."""
    if shape[0] % len(generators) != 0:
        raise ValueError("shape[0] needs to be a multiple of len(generators)")
    return torch.cat([torch.rand((1, *shape[1:]), generator=g, device=g.device, dtype=dtype)
                      for g in list(generators) * (shape[0] // len(generators))], dim=0).to(device)


class HiresUnetWrapper:
    """hires_fix.py:123-235: `latents` holds [lo (natural size, zero-padded into the full frame) ; hi]."""

    def __init__(self, unet_natural, unet_hires, generators, natural_size, oos_fraction, latent_debugger=None):
        self.unet_natural, self.unet_hires = unet_natural, unet_hires
        self.generators, self.natural_size, self.oos_fraction = generators, natural_size, oos_fraction
        self.easing = Easing(floor=0, start=0, end=0.667, easing="cubic")

    def __call__(self, latents, step, u):
        p = self.easing.interp(u)
        lo_in, hi_in = latents.chunk(2)
        if isinstance(step, torch.Tensor) and step.shape:
            lo_t, hi_t = step.chunk(2)
        else:
            lo_t = hi_t = step
        hi = self.unet_hires(hi_in, hi_t, u=u)
        if p >= 0.999:
            return torch.concat([lo_in, hi])
        *_, h, w = latents.shape
        th, tw = self.natural_size
        offseth, offsetw = (h - th) // 2, (w - tw) // 2
        lo_in = lo_in[:, :, offseth:offseth + th, offsetw:offsetw + tw]
        lo = self.unet_natural(lo_in, lo_t, u=u)
        hi_down = scale_into(hi, down_scale_factor(hi.shape, lo.shape, self.oos_fraction), target_shape=lo.shape)   # "pad"
        randmap = batched_rand(lo.shape, self.generators, lo.device, lo.dtype)
        lo_merged = torch.where(randmap >= p, lo, hi_down)
        lo_up = scale_into(lo, up_scale_factor(lo.shape, hi.shape, self.oos_fraction), target=hi.clone())           # "clone"
        randmap = batched_rand(hi.shape, self.generators, hi.device, hi.dtype)
        hi_merged = torch.where(randmap >= p, lo_up, hi)
        lo_expanded = torch.zeros_like(hi_merged)
        lo_expanded[:, :, offseth:offseth + th, offsetw:offsetw + tw] = lo_merged
        return torch.concat([lo_expanded, hi_merged])

    @classmethod
    def image_to_natural(cls, natural_size, image, oos_fraction):
        target_shape = [natural_size, natural_size]
        return scale_into(image, down_scale_factor(image.shape, target_shape, oos_fraction), target_shape=target_shape)

    @classmethod
    def merge_initial_latents(cls, left, right):
        left_resized = torch.zeros_like(right)
        *_, th, tw = left.shape
        *_, h, w = right.shape
        offseth, offsetw = (h - th) // 2, (w - tw) // 2
        left_resized[:, :, offseth:offseth + th, offsetw:offsetw + tw] = left
        return torch.concat([left_resized, right])

    @classmethod
    def split_result(cls, left, right):
        return right.chunk(2)[1]


class GraftUnets:
    """graft.py:16-56."""

    def __init__(self, unet_root, unet_top, generators, blend={}):
        self.unet_root, self.unet_top, self.generators = unet_root, unet_top, generators
        self.easing = Easing(**{"floor": 0, "start": 0.1, "end": 0.3, "easing": "sine", **blend})

    def __call__(self, latents, step, u):
        p = self.easing.interp(u)
        if p <= 0:
            return self.unet_root(latents, step, u=u)
        if p >= 1:
            return self.unet_top(latents, step, u=u)
        root = self.unet_root(latents, step, u=u)
        top = self.unet_top(latents, step, u=u)
        randmap = batched_rand(top.shape, self.generators, top.device, top.dtype)
        return torch.where(randmap >= p, root, top)

    @classmethod
    def merge_initial_latents(cls, left, right):
        return left

    @classmethod
    def split_result(cls, left, right):
        return right


# --------------------------------------------------------------------------------------------- pipeline composition
def hires_txt2img_latents(eps_unet_cfg, *, batch, height, width, sample_size, seeds, steps, oos_fraction=0.6,
                          latent_dtype=torch.float32, eta=1.0, prediction_type="epsilon", graft_top_cfg=None,
                          graft_blend=None):
    """txt2img with the hires fix engaged, the way UnifiedPipeline composes it (unified_pipeline.py:2100-2181 mode tree,
    :2461-2486 collapse / initial latents / loop / split_result) with the Euler-ancestral sampler: leaves share the CFG'd
    eps UNet, the natural leaf runs at sample_size^2, u = i / n per step (KDiffusionPositionTracker inside trange,
    common_scheduler.py:358-389).  With `graft_top_cfg` every leaf is a GraftUnets(root = eps_unet_cfg, top = graft_top_cfg)
    pair instead (graft + hires nesting)."""
    from . import sampling as S
    gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]

    def gen_latents(h, w):
        shape = (batch, 4, h, w)
        mid = S.batched_randn([batch, 4, sample_size, sample_size], gens, "cpu", latent_dtype)
        off2, off3 = (sample_size - h) // 2, (sample_size - w) // 2
        if off2 > 0:
            mid = mid[:, :, off2:off2 + h, :]
        if off3 > 0:
            mid = mid[:, :, :, off3:off3 + w]
        if off2 >= 0 and off3 >= 0:
            return mid
        lat = S.batched_randn(shape, gens, "cpu", latent_dtype)
        o2, o3 = (lat.shape[2] - mid.shape[2]) // 2, (lat.shape[3] - mid.shape[3]) // 2
        lat[:, :, o2:o2 + mid.shape[2], o3:o3 + mid.shape[3]] = mid
        return lat

    acp = S.sd_alphas_cumprod("cpu")
    Den = S.VDenoiser if prediction_type == "v_prediction" else S.EpsDenoiser
    den = Den(eps_unet_cfg, acp)
    sig_full = S.k_sigmas(den, steps)
    left = gen_latents(sample_size, sample_size) * sig_full[0]
    right = gen_latents(height // 8, width // 8) * sig_full[0]
    latents = HiresUnetWrapper.merge_initial_latents(left, right).float()
    sigmas = sig_full.to(latent_dtype).float()

    def leaf_of(d):
        return lambda x, sigma, u: d(x, sigma)
    if graft_top_cfg is None:
        nat = hi = leaf_of(den)
    else:
        top = Den(graft_top_cfg, acp)
        nat = GraftUnets(leaf_of(den), leaf_of(top), gens, blend=graft_blend or {})
        hi = GraftUnets(leaf_of(den), leaf_of(top), gens, blend=graft_blend or {})
    wrapper = HiresUnetWrapper(nat, hi, gens, [sample_size, sample_size], oos_fraction)
    n = len(sigmas) - 1
    calls = {"i": 0}

    def model(x, sigma):
        u = max(min(calls["i"] / n, 0.999), 0)
        calls["i"] += 1
        return wrapper(x, sigma, u)
    noise = lambda *_: S.batched_randn(latents.shape, gens, "cpu", latent_dtype).float()
    out = S.sample_euler_ancestral(model, latents, sigmas, noise, eta=eta)
    return HiresUnetWrapper.split_result(None, out)


def grafted_inpaint_latents(inpaint_unet, main_unet, vae, uncond_emb, cond_emb, guidance_scale, *, image, mask_image, seeds,
                            steps, strength, blend=None, latent_dtype=torch.float32):
    """Grafted inpaint as UnifiedPipeline composes it (unified_pipeline.py:2069-2098): GraftUnets(root = the inpaint UNet
    in EnhancedRunwayInpaintMode, top = the main UNet in EnhancedInpaintMode), both modes constructed and then both
    `generateLatents` run on the shared generators; the root's latents start the loop (graft.py:50-52)."""
    from . import sampling as S
    gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]
    kw = dict(image=image, mask_image=mask_image, generators=gens, steps=steps, strength=strength, latent_dtype=latent_dtype)
    ph_root = S.image_mode_phases(inpaint_unet, vae, uncond_emb, cond_emb, guidance_scale, **kw)
    ph_top = S.image_mode_phases(main_unet, vae, uncond_emb, cond_emb, guidance_scale, **kw)
    next(ph_root)
    next(ph_top)
    root = next(ph_root)
    top = next(ph_top)
    graft = GraftUnets(lambda x, s, u: root["k_unet"](x, s, u), lambda x, s, u: top["k_unet"](x, s, u), gens, blend=blend or {})
    return S.euler_ancestral_with_u(lambda x, s, u: graft(x, s, u), root["latents"], root["sigmas"], root["u_off"], gens,
                                    latent_dtype)


def hires_image_mode_latents(unet, vae, uncond_emb, cond_emb, guidance_scale, *, image, mask_image=None, seeds, steps,
                             strength, sample_size, oos_fraction=1.0, latent_dtype=torch.float32):
    """img2img / inpaint with the hires fix engaged (unified_pipeline.py:2100-2181): the natural leaf works on the image
    (and mask) shrunk by `image_to_natural` (oos_fraction defaults to `_hires_image_oos_fraction` = 1.0 when an image is
    given, :1840-1843); both leaves' modes are constructed, then both generate their latents, on shared generators."""
    from . import sampling as S
    gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]
    px = sample_size * 8
    nat_image = HiresUnetWrapper.image_to_natural(px, image, oos_fraction)
    nat_mask = None if mask_image is None else HiresUnetWrapper.image_to_natural(px, mask_image, oos_fraction)
    kw = dict(generators=gens, steps=steps, strength=strength, latent_dtype=latent_dtype)
    ph_n = S.image_mode_phases(unet, vae, uncond_emb, cond_emb, guidance_scale, image=nat_image, mask_image=nat_mask, **kw)
    ph_h = S.image_mode_phases(unet, vae, uncond_emb, cond_emb, guidance_scale, image=image, mask_image=mask_image, **kw)
    next(ph_n)
    next(ph_h)
    nat = next(ph_n)
    hi = next(ph_h)
    latents = HiresUnetWrapper.merge_initial_latents(nat["latents"], hi["latents"])
    w = HiresUnetWrapper(lambda x, s, u: nat["k_unet"](x, s, u), lambda x, s, u: hi["k_unet"](x, s, u), gens,
                         [sample_size, sample_size], oos_fraction)
    out = S.euler_ancestral_with_u(lambda x, s, u: w(x, s, u), latents, hi["sigmas"], hi["u_off"], gens, latent_dtype)
    return HiresUnetWrapper.split_result(None, out)


class WithExtraChannels:
    """UnetWithExtraChannels (gyre/pipeline/unet/core.py:21-37) around a DiffusersUNet-protocol object: the same un-scaled
    extra channels (a depth map) appended to the latents at every call."""

    def __init__(self, unet, extra):
        self.unet, self.extra, self.config = unet, extra, unet.config

    def __call__(self, latents, t, *, encoder_hidden_states, **kw):
        e = torch.cat([self.extra] * (latents.shape[0] // self.extra.shape[0]))
        return self.unet(torch.cat([latents, e.to(latents.dtype)], dim=1), t, encoder_hidden_states=encoder_hidden_states)


def depth_txt2img_latents(depth_unet, main_unet, uncond_emb, cond_emb, guidance_scale, *, depth_map, seeds, steps, sample_size,
                          height, width, graft_blend=None, latent_dtype=torch.float32):
    """A depth hint as UnifiedPipeline composes it (unified_pipeline.py:2004-2013, 2069-2098, 2305-2309): the request goes to
    the 5-channel depth UNet with the depth map (already 2 * d - 1 at latent resolution) as its fifth input channel; with
    `graft_blend` (engine option grafted_depth) GraftUnets(root = depth leaf, top = main UNet without the depth input) - both
    leaves draw their initial latents on the shared generators, the root's start the loop.  PINNED against
    UnifiedPipeline.__call__ (scripts/make_golden.py:pin_call)."""
    from . import sampling as S
    B = len(seeds)
    depth_cfg = S.CFGParallel(WithExtraChannels(depth_unet, depth_map.expand(B, -1, -1, -1).contiguous()), uncond_emb, cond_emb,
                              guidance_scale)
    if graft_blend is None:
        return S.txt2img_latents(depth_cfg, batch=B, in_channels=4, height=height, width=width, sample_size=sample_size,
                                 seeds=seeds, steps=steps, sampler="euler_a", latent_dtype=latent_dtype)
    gens = [torch.Generator("cpu").manual_seed(s) for s in seeds]
    acp = S.sd_alphas_cumprod()
    den_root = S.EpsDenoiser(depth_cfg, acp)
    den_top = S.EpsDenoiser(S.CFGParallel(main_unet, uncond_emb, cond_emb, guidance_scale), acp)
    sig = S.k_sigmas(den_root, steps)
    h, w = height // 8, width // 8
    lat_root = S.batched_randn([B, 4, h, w], gens, "cpu", latent_dtype) * sig[0]
    S.batched_randn([B, 4, h, w], gens, "cpu", latent_dtype)                  # the top leaf's (discarded) draw
    graft = GraftUnets(lambda x, s, u: den_root(x, s), lambda x, s, u: den_top(x, s), gens, blend=graft_blend)
    return S.euler_ancestral_with_u(lambda x, s, u: graft(x, s, u), lat_root, sig.float(), 0.0, gens, latent_dtype)


# ---------------------------------------------------------------------------------------------------------------------
# gyre/images.py:324-340 `resize(tensor, factors, sharpness)` for sharpness 1 / 2: ResizeRight with lanczos3, reflect padding,
# antialiasing iff sharpness == 1, clamped to [0, 1] - what hint masks (unified_pipeline.py:790-808 resized_mask) and the depth
# hint (:2007-2008) go through.  PINNED against the reference's own gyre/images.py + vendored ResizeRight
# (scripts/make_golden.py:pin_resize, tests/golden/resize.pt).
def lanczos3(x):
    """interp_methods.py:54-57."""
    eps = torch.finfo(torch.float32).eps
    return ((torch.sin(math.pi * x) * torch.sin(math.pi * x / 3) + eps) / ((math.pi ** 2 * x ** 2 / 3) + eps)) * (abs(x) < 3).to(x.dtype)


def resize_taps_general(in_sz: int, scale: float, support: float = 6.0, antialiasing: bool = True, method=lanczos3):
    """resize_right.py:70-118 for one dim: (source indices [out, K] with reflect padding folded in, weights [out, K], out_sz).
    Antialiasing (apply_antialiasing_if_needed :175-200) stretches the window by 1 / scale when downscaling."""
    eps = torch.finfo(torch.float32).eps
    scale = float(scale)
    out_sz = math.ceil(scale * in_sz)
    grid = torch.arange(out_sz) / scale + (in_sz - 1) / 2 - (out_sz - 1) / (2 * scale)
    if antialiasing and scale < 1.0:
        interp = lambda arg: scale * method(scale * arg)
        cur_support = support / scale
    else:
        interp, cur_support = method, support
    left = (grid - cur_support / 2 - eps).ceil().long()
    fov = left[:, None] + torch.arange(math.ceil(cur_support - eps))
    pad0 = -fov[0, 0].item()
    w = interp((grid + pad0)[:, None] - (fov + pad0))
    s = w.sum(1, keepdim=True)
    s[s == 0] = 1
    w = w / s
    idx = torch.where(fov < 0, -fov, fov)
    idx = torch.where(idx > in_sz - 1, 2 * (in_sz - 1) - idx, idx)
    assert idx.min() >= 0 and idx.max() <= in_sz - 1, "reflect padding wider than the image"
    return idx, w, out_sz


def images_resize(tensor, factors, sharpness=1):
    """gyre/images.py:324-340 (sharpness 1: antialiased, 2: not; 0 - the area-downscale variant - is not restated)."""
    if sharpness not in (1, 2):
        raise NotImplementedError("sharpness 0 (area downscale) is not restated")
    if not isinstance(factors, (tuple, list)):
        factors = (factors, factors)
    elif len(factors) == 1:
        factors = (factors[0], factors[0])
    out = tensor
    dims = sorted((-2, -1), key=lambda d: float(factors[d]))            # ascending scale, H before W on ties
    for d in dims:
        if float(factors[d]) == 1.0:
            continue
        idx, w, _ = resize_taps_general(out.shape[d], factors[d], antialiasing=sharpness == 1)
        t = out.transpose(d, 0)
        ww = w.reshape(*w.shape, *([1] * (t.ndim - 1)))
        out = (t[idx] * ww).sum(1).transpose(0, d)
    return out.to(tensor.dtype).clamp(0, 1)


def images_rescale(tensor, height, width=None, fit="cover", pad_mode="constant", sharpness=1):
    """gyre/images.py:369-408 (pinned with images_resize: scripts/make_golden.py:pin_resize)."""
    if width is None:
        width = height
    orig_h, orig_w = tensor.shape[-2], tensor.shape[-1]
    scale_h, scale_w = height / orig_h, width / orig_w
    if fit == "cover":
        scale_h = scale_w = max(scale_h, scale_w)
    elif fit == "contain":
        scale_h = scale_w = min(scale_h, scale_w)
    tensor = images_resize(tensor, (scale_h, scale_w), sharpness=sharpness)
    res_h, res_w = tensor.shape[-2], tensor.shape[-1]
    err_h, err_w = (height - res_h) // 2, (width - res_w) // 2
    top, left = (-err_h if err_h < 0 else 0), (-err_w if err_w < 0 else 0)
    tensor = tensor[:, :, top:top + height, left:left + width]
    pad = [err_w, width - res_w - err_w] if err_w > 0 else [0, 0]
    pad += [err_h, height - res_h - err_h] if err_h > 0 else [0, 0]
    return torch.nn.functional.pad(tensor, pad, pad_mode)
