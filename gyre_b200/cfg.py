"""Classifier-free-guidance wrapper around the native UNet: the reference's
`CFGUNet_Parallel` (gyre/pipeline/unet/cfg.py:41-57) composed with `UNetWithEmbeddings` /
`CFGUNetFromDiffusersUNet` (gyre/pipeline/unet/core.py:242-274): ONE UNet call on the doubled batch
`[uncond ; cond]`, then `u + s * (g - u)`."""
from __future__ import annotations

import torch

from . import _native as N


class B200GuidedUNet:
    """`NoisePredictionUNet` protocol object (`eps = f(latents, t)`, gyre/pipeline/unet/types.py:56-59) that
    also exposes the doubled-batch entry the fused scheduler loop uses (`raw`)."""

    def __init__(self, unet, uncond_embeddings, text_embeddings, guidance_scale: float, parallel: bool = True):
        if uncond_embeddings.shape != text_embeddings.shape:
            raise ValueError("uncond and text embeddings must have the same shape")
        self.unet = unet
        # the wrappers are built per request, after LoRA hooks were applied / removed (unified_pipeline.py:2193-2238):
        # fold them into the packed weights once here, not on every step
        sync = getattr(unet, "_sync_lora", None)
        if sync is not None:
            sync()
        self.guidance_scale = float(guidance_scale)
        self.batch = text_embeddings.shape[0]
        # CFGUNet_Parallel (one UNet call on the doubled batch, cfg.py:41-57) or CFGUNet_Sequential (two calls of the
        # single batch, cfg.py:26-38 - the reference's low-memory execution mode); same results, halves the workspace
        self.parallel = parallel
        # CFG order is [uncond, cond] (unified_pipeline.py:2335, cfg.py:54)
        self.embeddings = torch.cat([uncond_embeddings, text_embeddings]).to(device=unet.device,
                                                                              dtype=torch.float16).contiguous()
        # who "owns" the K/V projections cached in the UNet: leaves of one request that share the embeddings (hires-fix:
        # natural + full size) point this at the same object so that alternating between them does not re-project
        self.ctx_owner = self
        # `raw` is fed `torch.cat([latents] * 2)` with one timestep (CFGUNet_Parallel, cfg.py:47-57; the schedulers build
        # the same layout on the device): the native UNet computes what both halves share once.  Callers that hand `raw`
        # two DIFFERENT halves must clear this.
        self.duplicated_halves = True
        self.extra = None           # [B, Ce, h, w] fp16: mask + masked-image latents of the inpaint UNets
        self._xcat = None
        self.add_cond = None        # [2B, proj_in] fp16: text_time conditioning of SDXL-style UNets ([uncond ; cond])
        self.controlnets = []       # B200ControlnetHint objects (gyre_b200.hints): evaluated at every UNet call
        self.t2i = None             # UNetWithT2I-style provider of the per-request adapter states

    def set_added_cond(self, uncond_kwargs, cond_kwargs):
        """`added_cond_kwargs` of text_time models ({text_embeds [B, P], time_ids [B, 6]}), for the uncond and the
        cond half of the CFG batch."""
        te = torch.cat([uncond_kwargs["text_embeds"], cond_kwargs["text_embeds"]])
        ti = torch.cat([uncond_kwargs["time_ids"], cond_kwargs["time_ids"]])
        if te.shape[0] != 2 * self.batch:
            raise ValueError("added_cond_kwargs batch does not match the embeddings")
        self.add_cond = self.unet.added_cond_vector(te, ti)

    def set_hints(self, hints):
        """`hints` of a mode-tree leaf (unified_pipeline.py:2312-2323): grouped by kind like the reference groups them by
        class - the ControlNets run at every UNet call on the latents it is about to see, the T2I-adapter states are
        computed once here."""
        from .hints import B200ControlnetHint, B200T2iHint, UNetWithT2I
        hints = list(hints or [])
        unknown = [h for h in hints if not isinstance(h, (B200ControlnetHint, B200T2iHint))]
        if unknown:
            raise ValueError(f"unknown hint objects: {unknown}")
        self.controlnets = [h for h in hints if isinstance(h, B200ControlnetHint)]
        t2i = [h for h in hints if isinstance(h, B200T2iHint)]
        self.t2i = UNetWithT2I(None, t2i) if t2i else None
        if self.t2i is not None and self.t2i.style_states is not None:
            # style tokens extend the context for the life of this wrapper (core.py:221-237): [uncond + its own tail ; cond +
            # style tokens]; the UNet re-projects K / V for the longer context on first use
            self.embeddings = self.t2i.styled_context(self.embeddings, "f").to(torch.float16).contiguous()

    def _hint_kwargs(self, x_in, t_i64, embeddings, cfg_meta):
        """UNetWithControlnet / UNetWithT2I (unet/core.py:38-64, 212-239) for one native UNet call."""
        if not self.controlnets and self.t2i is None:
            return {}
        from .hints import controlnet_residual_kwargs
        kw = {}
        if self.controlnets:
            kw.update(controlnet_residual_kwargs(self.controlnets, x_in, t_i64, embeddings, cfg_meta))
            if not torch.is_tensor(kw["mid_block_additional_residual"]):        # every ControlNet was cfg_only on the "u" side
                kw = {}
        if self.t2i is not None:
            kw["adapter_states"] = self.t2i.states_for(cfg_meta, x_in.shape[0])
        return kw

    def set_extra_channels(self, extra):
        """EnhancedRunwayInpaintMode.wrap_unet (unified_pipeline.py:668-690) / UnetWithExtraChannels (unet/core.py:21-37):
        the same un-scaled extra channels are appended to the (CFG-doubled) latents at every step."""
        if extra is not None:
            N.require_cuda(extra)
            if extra.shape[0] != self.batch:
                raise ValueError("extra channels batch does not match the embeddings")
            if extra.shape[1] + 4 != self.unet.config.in_channels:
                raise ValueError(f"UNet takes {self.unet.config.in_channels} channels, got 4 + {extra.shape[1]}")
            extra = extra.to(torch.float16).contiguous()
        self.extra = extra

    def _expand(self, x2_f16):
        if self.extra is None:
            return x2_f16
        B2, Cx, h, w = x2_f16.shape
        Ce = self.extra.shape[1]
        if self._xcat is None or self._xcat.shape != (B2, Cx + Ce, h, w):
            self._xcat = torch.empty((B2, Cx + Ce, h, w), device=x2_f16.device, dtype=torch.float16)
        N.check(N.load().gyre_b200_cat_channels(N.ptr(x2_f16), Cx, N.ptr(self.extra), Ce, self.extra.shape[0], B2, h * w,
                                                N.ptr(self._xcat), N.stream_ptr(x2_f16.device)), "cat_channels")
        return self._xcat

    def raw(self, x2_f16, t2_i64, out=None):
        if not self.parallel:
            return self._raw_sequential(x2_f16, t2_i64, out)
        x2_f16 = self._expand(x2_f16)
        # The embeddings are fixed for the life of this wrapper (as in UNetWithEmbeddings, core.py:253-259), so
        # the UNet projects them to cross-attention K/V once instead of at every step (tunable CTX_KV_CACHE).
        if N.get_tunable("CTX_KV_CACHE"):
            bound = self.unet._ctx_bound
            if bound is None or bound[0] is not self.ctx_owner:
                self.unet.set_context(self.embeddings, owner=self.ctx_owner)
            return self.unet.forward_raw(x2_f16, t2_i64, None, out=out, add_cond=self.add_cond,
                                         cfg_duplicate=self.duplicated_halves,
                                         **self._hint_kwargs(x2_f16, t2_i64, self.embeddings, "f"))
        return self.unet.forward_raw(x2_f16, t2_i64, self.embeddings, out=out, add_cond=self.add_cond,
                                     cfg_duplicate=self.duplicated_halves,
                                     **self._hint_kwargs(x2_f16, t2_i64, self.embeddings, "f"))

    def _raw_sequential(self, x2_f16, t2_i64, out):
        """[uncond ; cond] evaluated as two UNet calls of batch B (CFGUNet_Sequential)."""
        B = self.batch
        if out is None:
            out = torch.empty((2 * B, self.unet.config.out_channels, *x2_f16.shape[2:]), device=x2_f16.device,
                              dtype=torch.float16)
        xe = self._expand(x2_f16)
        for half in (0, 1):
            sl = slice(half * B, (half + 1) * B)
            add = self.add_cond[sl].contiguous() if self.add_cond is not None else None
            xh, th, eh = xe[sl].contiguous(), t2_i64[sl].contiguous(), self.embeddings[sl].contiguous()
            self.unet.forward_raw(xh, th, eh, out=out[sl], add_cond=add, **self._hint_kwargs(xh, th, eh, "ug"[half]))
        return out

    def __call__(self, latents, t):
        N.require_cuda(latents)
        B = latents.shape[0]
        if B != self.batch:
            raise ValueError(f"latents batch {B} does not match the {self.batch} bound embeddings")
        x = latents.to(torch.float16)
        x2 = torch.cat([x, x]).contiguous()
        t2 = self.unet._timesteps(t, B)
        t2 = torch.cat([t2, t2])
        eps2 = self.raw(x2, t2)
        out16 = torch.empty_like(x) if latents.dtype == torch.float16 else None
        out32 = torch.empty(latents.shape, device=latents.device, dtype=torch.float32) if out16 is None else None
        N.check(N.load().gyre_b200_cfg_combine(N.ptr(eps2), self.guidance_scale, B, x[0].numel(), N.ptr(out16),
                                               N.ptr(out32), N.stream_ptr(latents.device)), "cfg_combine")
        return out16 if out16 is not None else out32.to(latents.dtype)
