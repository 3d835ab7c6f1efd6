"""Graft of two scheduler UNets (reference: gyre/pipeline/unet/graft.py:16-56; used for grafted inpaint / depth,
unified_pipeline.py:2069-2098): below the eased window only the root leaf runs, above it only the top leaf, inside it
both predict x0 and a per-pixel uniform map picks between them - one `gyre_b200_rand_select` launch."""
from __future__ import annotations

import torch

from . import _native as N
from .easing import Easing
from .hires_fix import _threshold, batched_rand


class GraftUnets:
    def __init__(self, unet_root, unet_top, generators, blend={}, rand_dtype=torch.float16):
        self.unet_root = unet_root
        self.unet_top = unet_top
        self.generators = generators
        self.easing = Easing(**{"floor": 0, "start": 0.1, "end": 0.3, "easing": "sine", **blend})
        self.rand_dtype = rand_dtype

    def __call__(self, latents, step, u: float):
        p = self.easing.interp(u)
        if p <= 0:
            return self.unet_root(latents, step, u=u)
        if p >= 1:
            return self.unet_top(latents, step, u=u)
        root = self.unet_root(latents, step, u=u).contiguous()
        top = self.unet_top(latents, step, u=u).contiguous()
        N.require_cuda(root, top)
        randmap = batched_rand(top.shape, self.generators, top.device, self.rand_dtype).float().contiguous()
        out = torch.empty_like(top)
        N.check(N.load().gyre_b200_rand_select(N.ptr(root), N.ptr(top), N.ptr(randmap), _threshold(p, self.rand_dtype),
                                               top.numel(), N.ptr(out), N.stream_ptr(top.device)), "rand_select")
        return out

    @classmethod
    def merge_initial_latents(cls, left, right):
        return left

    @classmethod
    def split_result(cls, left, right):
        return right
