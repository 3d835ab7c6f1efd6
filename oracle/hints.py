"""Oracle: ControlNet / T2I-adapter hints between the CFG wrapper and the UNet, restated on the CPU (TEST ONLY).

  * `UNetWithControlnet`, `AdapterStateList`, `UNetWithT2I` follow gyre/pipeline/unet/core.py:38-64, 67-94, 97-239 (standard
    adapter states only; no co-adapters / style states).  core.py is pure torch and importable: PINNED by
    scripts/make_golden.py:pin_hints against the reference's own classes over fake models (tests/golden/hints.pt).
  * `ControlnetHint.__call__`, `T2iHint.__call__` follow gyre/pipeline/unified_pipeline.py:957-1058 and :925-939
    (UnifiedPipelineHint_Controlnet / _T2i.standard_call).  unified_pipeline.py imports diffusers at module level and cannot
    be loaded plainly; scripts/_vendored.py:gyre_unified_pipeline loads it with the absent third-party packages stood in for
    by empty classes, and scripts/make_golden.py:pin_hint_classes runs the reference's two classes inside the reference's own
    wrapper stack / scheduler / Txt2imgMode: PINNED, 7 runs bit-identical (tests/golden/hint_classes.pt; mask-less hints,
    4-channel latents)."""
from __future__ import annotations

from types import SimpleNamespace

import torch


def split_hint(image, mask, channels=3):
    """UnifiedPipelineHint.__init__ (unified_pipeline.py:748-775)."""
    if mask is None and image.shape[1] == 4:
        mask = image[:, [3]]
    image = image[:, [0]] if channels == 1 else (image[:, [0, 1, 2]] if image.shape[1] >= 3 else image[:, [0, 0, 0]])
    if mask is not None:
        mask = mask[:, [0]]
        if mask.mean() == 1.0 and mask.std() == 0.0:
            mask = None
    return image, mask


def resized_mask(state, mask):
    """unified_pipeline.py:790-808."""
    from .hires import images_resize
    if mask is None:
        return torch.ones_like(state)
    hd, wd = (mask.shape[-2], state.shape[-2]), (mask.shape[-1], state.shape[-1])
    assert max(hd) % min(hd) == 0 and max(wd) % min(wd) == 0
    return images_resize(mask, (state.shape[-2] / mask.shape[-2], state.shape[-1] / mask.shape[-1])).to(state.dtype)


class ControlnetHint:
    def __init__(self, model, image, weight=1.0, soft_injection=False, cfg_only=False, mask=None):
        """model(cnlatents, t, encoder_hidden_states=, controlnet_cond=) -> object with down_block_res_samples, mid_block_res_sample."""
        self.model = model
        self.image, self.mask = split_hint(image, mask)
        self.weight, self.soft_injection, self.cfg_only = weight, soft_injection, cfg_only

    def __call__(self, latents, t, encoder_hidden_states, cfg_meta=None):
        cnlatents = latents[:, 0:4]
        if self.cfg_only:
            if cfg_meta == "f":
                cnlatents = cnlatents.chunk(2)[-1]
                t = t.chunk(2)[-1] if isinstance(t, torch.Tensor) else t
                encoder_hidden_states = encoder_hidden_states.chunk(2)[-1]
            elif cfg_meta == "u":
                return SimpleNamespace(down_block_res_samples=[torch.tensor(0)] * 13, mid_block_res_sample=torch.tensor(0))
        condition, mask = self.image, self.mask
        if latents.shape[1] == 9:
            mask = latents[:, [4]]
            cnlatents = cnlatents * mask
            mask = resized_mask(condition, mask)
            if self.mask is not None:
                mask = mask * self.mask
            condition = condition * mask
        res = self.model(cnlatents, t, encoder_hidden_states=encoder_hidden_states, controlnet_cond=condition)

        def basic(r, lw):
            return r * self.weight * lw * resized_mask(r, mask)

        process = basic
        if self.cfg_only and cfg_meta == "f":
            process = lambda r, lw: torch.cat([torch.zeros_like(r), basic(r, lw)], dim=0)
        layer_weights = torch.logspace(-1, 0, 13) if self.soft_injection else [1] * 13
        down = [process(r, lw) for r, lw in zip(res.down_block_res_samples, layer_weights)]
        mid = process(res.mid_block_res_sample, layer_weights[-1])
        return SimpleNamespace(down_block_res_samples=down, mid_block_res_sample=mid)


class T2iHint:
    def __init__(self, model, image, weight=1.0, soft_injection=False, cfg_only=False, mask=None, channels=3, style=None):
        """style: None for standard adapters, else dict(vision=callable(image, output_hidden_states) -> object with
        last_hidden_state / hidden_states, mean, std, size, clip_layer) - UnifiedPipelineHint_T2i.style_setup / style_call
        (unified_pipeline.py:843-853, 941-975)."""
        self.model = model
        self.image, self.mask = split_hint(image, None if style else mask, channels)
        self.weight, self.soft_injection, self.cfg_only = weight, (False if style else soft_injection), cfg_only
        self.fuser = None
        self.style = style

    def coadapter_type(self):
        return False

    def style_call(self):
        from .hires import images_rescale
        st = self.style
        layer = st.get("clip_layer") or "final"
        if isinstance(layer, int):
            layer = abs(layer)
        layer = "final" if layer in (0, 1, "final") else ("penultimate" if layer in (2, "penultimate") else layer)
        image = images_rescale(self.image, st["size"], st["size"], "cover")
        mean = torch.tensor(st["mean"]).view(1, 3, 1, 1)
        std = torch.tensor(st["std"]).view(1, 3, 1, 1)
        image = (image - mean) / std
        out = st["vision"](image, output_hidden_states=(layer != "final"))
        hidden = out.last_hidden_state if layer == "final" else (out.hidden_states[-2] if layer == "penultimate" else out.hidden_states[-layer])
        return self.model(hidden) * self.weight

    def __call__(self):
        if self.style:
            return self.style_call()
        layer_weights = (1.0, 1.0, 1.0, 1.0)
        if self.soft_injection:
            layer_weights = torch.logspace(-0.25, 0, 4)
            if self.cfg_only:
                layer_weights[0] = 0.25
        return [state * self.weight * lw * resized_mask(state, self.mask) for state, lw in zip(self.model(self.image), layer_weights)]


class UNetWithControlnet:
    def __init__(self, unet, controlnets):
        self.unet, self.controlnets = unet, controlnets

    def __call__(self, latents, t, **kwargs):
        cnargs = {"encoder_hidden_states": kwargs.get("encoder_hidden_states"), "cfg_meta": kwargs.get("cfg_meta")}
        residuals = [cn(latents, t, **cnargs) for cn in self.controlnets]
        resargs = {"down_block_additional_residuals": [sum(i) for i in zip(*[r.down_block_res_samples for r in residuals])],
                   "mid_block_additional_residual": sum([r.mid_block_res_sample for r in residuals])}
        return self.unet(latents, t, **kwargs, **resargs)


class AdapterStateList:
    def __init__(self):
        self.items = []

    def append(self, item, cfg_only):
        self.items.append((item, cfg_only))

    def _zeros_like(self, item):
        return torch.zeros_like(item) if isinstance(item, torch.Tensor) else [self._zeros_like(s) for s in item]

    @property
    def all(self):
        for item, _ in self.items:
            yield item

    @property
    def cfg_only(self):
        for item, cfg_only in self.items:
            yield item if cfg_only else self._zeros_like(item)

    @property
    def either(self):
        for item, cfg_only in self.items:
            yield item if not cfg_only else self._zeros_like(item)


class UNetWithT2I:
    def __init__(self, unet, t2i_adapters):
        self.unet = unet
        self.standard_states = None
        self.style_states = None
        standard = AdapterStateList()
        style = []
        for adapter in t2i_adapters:
            assert not adapter.coadapter_type()
            state = adapter()
            if isinstance(state, list):
                standard.append(state, adapter.cfg_only)
            else:
                style.append(state)
        if style:
            self.style_states = torch.cat(style, dim=1)
            self.style_dim0, self.style_dim1 = self.style_states.shape[0], self.style_states.shape[1]
        g, u = list(standard.all), list(standard.either)
        if g:
            self.standard_states = {"u": [sum(i) for i in zip(*u)], "g": [sum(i) for i in zip(*g)]}
            self.standard_dim0 = self.standard_states["g"][0].shape[0]
            self.standard_states["f"] = [torch.cat([a, b], dim=0) for a, b in zip(self.standard_states["u"], self.standard_states["g"])]

    def __call__(self, latents, t, **kwargs):
        is_f = kwargs["encoder_hidden_states"].shape[0] == self.standard_dim0 * 2
        cfg_meta = kwargs.get("cfg_meta", "f" if is_f else "g")
        if self.standard_states is not None:
            kwargs["adapter_states"] = self.standard_states[cfg_meta]
        if self.style_states is not None:                      # core.py:221-237
            hidden_states = kwargs.pop("encoder_hidden_states")
            if cfg_meta == "f":
                uncond, cond = hidden_states.chunk(2)
            elif cfg_meta == "u":
                uncond, cond = hidden_states, None
            else:
                uncond, cond = None, hidden_states
            res = []
            if uncond is not None:
                res += [torch.cat([uncond, uncond[:, -self.style_dim1:, :]], dim=1)]
            if cond is not None:
                res += [torch.cat([cond, self.style_states], dim=1)]
            kwargs["encoder_hidden_states"] = torch.cat(res, dim=0)
        return self.unet(latents, t, **kwargs)


class FromDiffusersUNet:
    """CFGUNetFromDiffusersUNet (core.py:262-274): drops cfg_meta, unwraps .sample."""

    def __init__(self, unet):
        self.unet = unet

    def __call__(self, latents, t, **kwargs):
        kwargs.pop("cfg_meta", None)
        return self.unet(latents, t, **kwargs).sample


class UNetWithEmbeddings:
    """core.py:242-259."""

    def __init__(self, unet, text_embeddings, cfg_meta):
        self.unet, self.text_embeddings, self.cfg_meta = unet, text_embeddings, cfg_meta

    def __call__(self, latents, t):
        return self.unet(latents, t, encoder_hidden_states=self.text_embeddings, cfg_meta=self.cfg_meta)


def guided_eps_unet(diffusers_unet, uncond, cond, guidance_scale, hints, parallel=True):
    """The wrapper stack of unified_pipeline.py:2284-2337 for one leaf: hints grouped by class (T2I / ControlNet wrap in the
    order they were first seen), embeddings bound, CFG parallel ("f") or sequential ("u", "g")."""
    unet = FromDiffusersUNet(diffusers_unet)
    grouped = {}
    for h in hints:
        grouped.setdefault(type(h), []).append(h)
    for cls, hs in grouped.items():
        unet = (UNetWithControlnet if cls is ControlnetHint else UNetWithT2I)(unet, hs)
    if parallel:
        f = UNetWithEmbeddings(unet, torch.cat([uncond, cond]), "f")

        def eps(latents, t):
            t2 = torch.cat([t, t]) if isinstance(t, torch.Tensor) and t.shape else t
            u, g = f(torch.cat([latents, latents]), t2).chunk(2)
            return u + guidance_scale * (g - u)
        return eps
    uu, gg = UNetWithEmbeddings(unet, uncond, "u"), UNetWithEmbeddings(unet, cond, "g")

    def eps(latents, t):
        u, g = uu(latents, t), gg(latents, t)
        return u + guidance_scale * (g - u)
    return eps
