"""Hires-fix scheduler-UNet wrapper on the native kernels (reference: gyre/pipeline/unet/hires_fix.py:95-235; engaged by
default for requests above the UNet's native size, unified_pipeline.py:1368-1370, 2100-2181).

The latents of a hires run hold [lo ; hi]: the natural-size sample (zero-padded into the full frame) and the full-size
sample.  Every step both leaves predict x0; while the eased progress p < 1 the two predictions are cross-blended through
per-pixel uniform maps: hi is lanczos-downscaled into lo's frame, lo is upscaled into hi's.  The reference does that
with ~20 torch ops per step; here each direction is ONE launch of `gyre_b200_resample_select` (resample + placement +
`where(rand >= p, ..)` + zero-frame insertion).  The per-dimension tap tables are computed on the host with
ResizeRight's fp32 expressions (a few hundred floats, cached per shape), so the interpolation weights are the
reference's bit for bit; the uniform maps are the reference's own `batched_rand` draws (seed contract).
"""
from __future__ import annotations

import math

import torch

from . import _native as N
from .easing import Easing

Hi, Wi = -2, -1


def down_scale_factor(latents_shape, target_shape, oos_fraction):
    """hires_fix.py:95-99: blend between fitting the short side (everything inside) and the long side (fill)."""
    scales = target_shape[Hi] / latents_shape[Hi], target_shape[Wi] / latents_shape[Wi]
    return min(*scales) * oos_fraction + max(*scales) * (1 - oos_fraction)


def up_scale_factor(latents_shape, target_shape, oos_fraction):
    return 1 / down_scale_factor(target_shape, latents_shape, oos_fraction)


def batched_rand(shape, generators, device, dtype):
    """gyre/pipeline/randtools.py:11-36: one `torch.rand((1, *shape[1:]))` per generator, on the generator's device."""
    if shape[0] % len(generators) != 0:
        raise ValueError(f"shape[0] ({shape[0]}) needs to be a multiple of len(generators) ({len(generators)})")
    draws = [torch.rand((1, *shape[1:]), generator=g, device=g.device, dtype=dtype)
             for g in list(generators) * (shape[0] // len(generators))]
    return torch.cat(draws, dim=0).to(device)


_TAPS = {}


def lanczos2_taps(in_sz: int, scale: float, device):
    """1-D plan of ResizeRight for one dimension (resize_right.py:128-213 with interp_methods.lanczos2, support 4,
    antialiasing off, replicate padding): out_sz = ceil(scale * in_sz); for every output position 4 clamped source indices
    (int32 [out, 4]) and normalised fp32 weights ([out, 4]).  Returns (idx, w, out_sz), cached per (in_sz, scale, device)."""
    key = (in_sz, float(scale), str(device))
    hit = _TAPS.get(key)
    if hit is not None:
        return hit
    eps = torch.finfo(torch.float32).eps
    out_sz = math.ceil(scale * in_sz)
    grid = torch.arange(out_sz) / float(scale) + (in_sz - 1) / 2 - (out_sz - 1) / (2 * float(scale))
    left = (grid - 4 / 2 - eps).ceil().long()
    fov = left[:, None] + torch.arange(math.ceil(4 - eps))
    pad0 = -fov[0, 0].item()                 # the generalised left pad shifts grid and field of view alike (:157-172)
    x = (grid + pad0)[:, None] - (fov + pad0)
    w = ((torch.sin(math.pi * x) * torch.sin(math.pi * x / 2) + eps) / ((math.pi ** 2 * x ** 2 / 2) + eps)) * (abs(x) < 2).to(x.dtype)
    tot = w.sum(1, keepdim=True)
    tot[tot == 0] = 1
    w = (w / tot).to(torch.float32)
    hit = (fov.clamp(0, in_sz - 1).to(torch.int32).contiguous().to(device), w.contiguous().to(device), out_sz)
    if len(_TAPS) > 64:
        _TAPS.clear()
    _TAPS[key] = hit
    return hit


def scale_into(src, scale, *, target=None, target_shape=None, other=None, rand_map=None, p=0.0, resampled_if_ge=True,
               frame=None):
    """`scale_into` (hires_fix.py:45-92, lanczos mode) as one launch, optionally fused with the blend that follows it.

    src [B, C, h, w] fp32 on the device is resized by `scale` and placed in the middle of a frame of `target_shape`
    (replicate-padded, strategy "pad") or of `target` (which provides the pixels outside, strategy "clone"; `target` is not
    modified).  With `other` / `rand_map` the result is where(rand_map >= p, resized, other) (or the two swapped);
    with `frame` = (FH, FW, oy, ox) it is written into a zero frame at that offset."""
    if (target is None) == (target_shape is None):
        raise ValueError("Only exactly one of target or target_shape")
    N.require_cuda(src)
    src = src.contiguous()
    B, C, SH, SW = src.shape
    tshape = tuple(target.shape if target is not None else target_shape)
    TH, TW = tshape[Hi], tshape[Wi]
    dev = src.device
    if float(scale) == 1.0:                      # ResizeRight skips dims whose scale is exactly 1
        ty = tx = (None, None, None)
        RH, RW = SH, SW
    else:
        ty = lanczos2_taps(SH, scale, dev)
        tx = lanczos2_taps(SW, scale, dev)
        RH, RW = ty[2], tx[2]
    offy, offx = (TH - RH) // 2, (TW - RW) // 2   # floor division: negative = centre crop starting at -off
    FH, FW, oy, ox = frame if frame is not None else (TH, TW, 0, 0)
    out = torch.empty((B, C, FH, FW), device=dev, dtype=torch.float32)
    for t in (target, other, rand_map):
        if t is not None and (tuple(t.shape) != (B, C, TH, TW) or t.dtype != torch.float32 or not t.is_contiguous()):
            raise ValueError("scale_into: target / other / rand_map must be contiguous fp32 tensors of the target shape")
    N.check(N.load().gyre_b200_resample_select(
        N.ptr(src), B * C, SH, SW, N.ptr(ty[0]), N.ptr(ty[1]), RH, N.ptr(tx[0]), N.ptr(tx[1]), RW, TH, TW, offy, offx,
        1 if target is not None else 0, N.ptr(target), N.ptr(other), N.ptr(rand_map), float(p), 1 if resampled_if_ge else 0,
        N.ptr(out), FH, FW, oy, ox, N.stream_ptr(dev)), "resample_select")
    return out


def _threshold(p: float, dtype) -> float:
    """`randmap >= p` compares in the map's dtype: a python scalar is cast to it first."""
    return float(torch.tensor(p, dtype=dtype))


class HiresUnetWrapper:
    """`GenericSchedulerUNet` composing two k-unet leaves (hires_fix.py:123-205).  Same constructor as the reference;
    `latents` are the fp32 device latents of the B200 schedulers."""

    def __init__(self, unet_natural, unet_hires, generators, natural_size, oos_fraction, latent_debugger=None,
                 rand_dtype=torch.float16):
        self.unet_natural = unet_natural
        self.unet_hires = unet_hires
        self.generators = generators
        self.natural_size = natural_size
        self.oos_fraction = oos_fraction
        self.easing = Easing(floor=0, start=0, end=0.667, easing="cubic")
        self.latent_debugger = latent_debugger
        self.rand_dtype = rand_dtype            # the reference draws the maps in the latents' dtype

    def __call__(self, latents, step, u: float):
        p = self.easing.interp(u)
        lo_in, hi_in = latents.chunk(2)
        if isinstance(step, torch.Tensor) and step.shape:
            lo_t, hi_t = step.chunk(2)
        else:
            lo_t = hi_t = step
        hi = self.unet_hires(hi_in.contiguous(), hi_t, u=u)
        if p >= 0.999:                           # past the graft stage: the natural-size half is carried along untouched
            return torch.concat([lo_in, hi])
        *_, h, w = latents.shape
        th, tw = self.natural_size
        offseth, offsetw = (h - th) // 2, (w - tw) // 2
        lo = self.unet_natural(lo_in[:, :, offseth:offseth + th, offsetw:offsetw + tw].contiguous(), lo_t, u=u)
        dev = latents.device
        pt = _threshold(p, self.rand_dtype)
        # lo <- where(rand >= p, lo, downscaled hi), written into the zero frame (`lo_expanded`)
        rand_lo = batched_rand(lo.shape, self.generators, dev, self.rand_dtype).float()
        lo_expanded = scale_into(hi, down_scale_factor(hi.shape, lo.shape, self.oos_fraction), target_shape=lo.shape,
                                 other=lo, rand_map=rand_lo, p=pt, resampled_if_ge=False, frame=(h, w, offseth, offsetw))
        # hi <- where(rand >= p, upscaled lo over a clone of hi, hi)
        rand_hi = batched_rand(hi.shape, self.generators, dev, self.rand_dtype).float()
        hi_merged = scale_into(lo, up_scale_factor(lo.shape, hi.shape, self.oos_fraction), target=hi, other=hi,
                               rand_map=rand_hi, p=pt, resampled_if_ge=True)
        return torch.concat([lo_expanded, hi_merged])

    @classmethod
    def image_to_natural(cls, natural_size: int, image, oos_fraction: float, fill=None):
        """hires_fix.py:207-218: the request's image / mask shrunk into the natural-size square (pixels, on the device)."""
        target_shape = [natural_size, natural_size]
        out = scale_into(image.float(), down_scale_factor(image.shape, target_shape, oos_fraction),
                         target_shape=(*image.shape[:2], *target_shape))
        return out.to(image.dtype)

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
