"""Hint images (ControlNet / T2I-adapter) between the CFG wrapper and the native UNet (the callers' side).

Mirrors, with the reference's names and argument meaning:
  * `UnifiedPipelineHint_Controlnet.__call__`  (gyre/pipeline/unified_pipeline.py:957-1058): which half of a CFG batch the
    ControlNet sees (`cfg_only`), per-layer `soft_injection` weights `logspace(-1, 0, 13)`, `weight`;
  * `UnifiedPipelineHint_T2i.standard_call`    (:925-939): `state * weight * layer_weight`, `logspace(-0.25, 0, 4)` with the
    first entry 0.25 for cfg_only;
  * `UNetWithControlnet`, `AdapterStateList`, `UNetWithT2I` (gyre/pipeline/unet/core.py:38-64, 67-94, 97-239): the sums over
    several ControlNets / adapters and the u / g / f (unconditional / guided / fused) selection.
The models are `B200ControlNet` / `B200T2iAdapter`; their outputs go to the native UNet through
`gyre_b200_unet_set_control_residuals` / `_set_adapter_states` (B200UNet.forward_raw keywords).

Hint masks (the alpha channel of an RGBA hint, or `mask=`) and ControlNets under the 9-channel inpaint UNets go through
`gyre_b200.images.resize` (the lanczos3 ResizeRight call of gyre/images.py:324-340 on the device); the resized masks are
cached per request - they do not change between steps.  Style adapters (`style_call`, :941-975: the hint image through
`images.rescale(.., "cover")`, the CLIP normalisation, the CLIP vision tower and `B200T2iStyleAdapter`) add context tokens to
the guided side (core.py:221-237).  Not built (raise): co-adapters + fuser."""
from __future__ import annotations

from types import SimpleNamespace

import torch

CONTROLNET_LAYERS = 13


def _split_hint(image, mask, channels=3):
    """UnifiedPipelineHint.__init__ (unified_pipeline.py:748-775): an RGBA hint carries its mask in the alpha channel; a mask
    of pure ones is dropped; image -> `channels` channels, mask -> 1."""
    if image.ndim == 3:
        image = image[None]
    if mask is None and image.shape[1] == 4:
        mask = image[:, [3]]
    if channels == 1:
        image = image[:, [0]]
    else:
        image = image[:, [0, 1, 2]] if image.shape[1] >= 3 else image[:, [0, 0, 0]]
    if mask is not None:
        mask = mask[None] if mask.ndim == 3 else mask
        mask = mask[:, [0]]
        if float(mask.float().mean()) == 1.0 and float(mask.float().std()) == 0.0:
            mask = None
    return image, mask


class _MaskResizer:
    """`resized_mask` (unified_pipeline.py:790-808): the mask scaled to a state's resolution with images.resize; the two sizes
    must be integer multiples of each other.  Cached per (mask, target size)."""

    def __init__(self):
        self._cache = {}

    def __call__(self, state, mask):
        from .images import resize
        key = (mask.data_ptr(), tuple(mask.shape), tuple(state.shape[-2:]))
        hit = self._cache.get(key)
        if hit is None:
            hd, wd = (mask.shape[-2], state.shape[-2]), (mask.shape[-1], state.shape[-1])
            if max(hd) % min(hd) or max(wd) % min(wd):
                raise ValueError(f"hint mask {tuple(mask.shape[-2:])} is not an integer multiple of the state {tuple(state.shape[-2:])}")
            scale = (state.shape[-2] / mask.shape[-2], state.shape[-1] / mask.shape[-1])
            hit = (mask, resize(mask.float(), scale).to(state.dtype))          # (keeps `mask` alive: the key is its address)
            if len(self._cache) > 64:
                self._cache.clear()
            self._cache[key] = hit
        return hit[1]


class B200ControlnetHint:
    """UnifiedPipelineHint_Controlnet (unified_pipeline.py:957-1058)."""

    def __init__(self, model, image, mask=None, weight=1.0, soft_injection=False, cfg_only=False):
        image, mask = _split_hint(image, mask)
        self.model = model
        self.image = image.to(device=model.device, dtype=torch.float16).contiguous()
        self.mask = None if mask is None else mask.to(device=model.device, dtype=torch.float16).contiguous()
        self._resized = _MaskResizer()
        self._masked_cond = {}
        self.weight = float(weight)
        self.soft_injection = bool(soft_injection)
        self.cfg_only = bool(cfg_only)
        self._cond = None
        # soft injection: later layers count more (the reference's experimentally chosen guess-mode weights)
        lw = torch.logspace(-1, 0, CONTROLNET_LAYERS).tolist() if self.soft_injection else [1.0] * CONTROLNET_LAYERS
        self.layer_weights = lw

    def to(self, device=None, dtype=None):
        return self

    def _condition(self, B):
        if self._cond is None or self._cond.shape[0] != B:
            if self.image.shape[0] not in (1, B):
                raise ValueError(f"hint image batch {self.image.shape[0]} does not match the UNet batch {B}")
            self._cond = self.image.expand(B, -1, -1, -1).contiguous()
        return self._cond

    def __call__(self, latents, t, encoder_hidden_states, cfg_meta=None):
        cnlatents = latents[:, 0:4]
        B_full = latents.shape[0]
        mask = self.mask
        fused_cfg_only = self.cfg_only and cfg_meta == "f"
        if self.cfg_only:
            if cfg_meta == "f":
                # only the guided half goes through the ControlNet; the unconditional half gets zeros
                cnlatents = cnlatents.chunk(2)[-1]
                t = t.chunk(2)[-1] if isinstance(t, torch.Tensor) and t.ndim > 0 else t
                encoder_hidden_states = encoder_hidden_states.chunk(2)[-1]
            elif cfg_meta == "u":
                return SimpleNamespace(down_block_res_samples=[0] * CONTROLNET_LAYERS, mid_block_res_sample=0)
        B = cnlatents.shape[0]
        condition = self._condition(B)
        if latents.shape[1] == 9:
            # under an inpaint UNet (:1008-1018) the ControlNet sees the latents and its conditioning image through the inpaint
            # mask (channel 4 of the UNet input, the same at every step), scaled up to the image with images.resize
            if self.cfg_only and cfg_meta == "f":
                raise NotImplementedError("cfg_only ControlNet hints under a 9-channel UNet in one fused CFG batch "
                                          "(the reference's own shapes do not line up there)")
            lmask = latents[:, [4]]
            cnlatents = cnlatents * lmask
            key = (lmask.data_ptr(), tuple(lmask.shape))
            hit = self._masked_cond.get(key)
            if hit is None:
                m = self._resized(condition, lmask.contiguous())
                if self.mask is not None:
                    m = m * self.mask
                hit = (m, (condition * m).contiguous())
                self._masked_cond = {key: hit}
            mask, condition = hit
        out = None
        if fused_cfg_only:
            # [zeros ; residual] without a concatenation per tensor per step: the ControlNet writes the second half
            shapes, mid_shape = self.model._skip_shapes(B_full, *latents.shape[2:])
            full = [torch.zeros(s, device=latents.device, dtype=torch.float16) for s in shapes + [mid_shape]]
            out = [f[B_full // 2:] for f in full]
        res = self.model(cnlatents.contiguous(), t, encoder_hidden_states=encoder_hidden_states.contiguous(),
                         controlnet_cond=condition, out=out)
        down = list(full[:-1]) if fused_cfg_only else list(res.down_block_res_samples)
        mid = full[-1] if fused_cfg_only else res.mid_block_res_sample
        n = len(down)
        lw = self.layer_weights if n == CONTROLNET_LAYERS - 1 else [1.0] * (n + 1)   # 12 skips + mid = the 13 layers
        for k in range(n):
            s = self.weight * lw[k]
            if s != 1.0:
                down[k].mul_(s)
            if mask is not None:
                down[k].mul_(self._resized(down[k], mask))
        s = self.weight * lw[-1]
        if s != 1.0:
            mid.mul_(s)
        if mask is not None:
            mid.mul_(self._resized(mid, mask))
        return SimpleNamespace(down_block_res_samples=down, mid_block_res_sample=mid)


def normalise_clip_layer(clip_layer, default=None):
    """gyre/pipeline/prompt_types.py:19-34."""
    if clip_layer is None:
        clip_layer = default
    if isinstance(clip_layer, int):
        clip_layer = abs(clip_layer)
    if clip_layer in (0, 1, "final"):
        return "final"
    if clip_layer in (2, "penultimate"):
        return "penultimate"
    return clip_layer


class B200T2iHint:
    """UnifiedPipelineHint_T2i (unified_pipeline.py:836-955): standard adapters (a list of per-level states) and style
    adapters (context tokens; needs `clip_model` = B200CLIPVisionModel and a feature extractor for mean / std / size)."""

    def __init__(self, model, image, mask=None, weight=1.0, soft_injection=False, cfg_only=False, clip_model=None,
                 feature_extractor=None, clip_layer=None):
        self.model = model
        self.type = "style" if hasattr(model, "num_token") else "standard"
        if self.type == "style":
            # (masks and soft injection "don't make sense" for style adapters: ignored like the reference, :873-883)
            if clip_model is None or feature_extractor is None:
                raise ValueError("T2i Style model needs a clip model and feature extractor")
            img, _ = _split_hint(image, None)
            self.image = img.to(device=model.device, dtype=torch.float32).contiguous()
            self.mask = None
            self.weight, self.soft_injection, self.cfg_only = float(weight), False, bool(cfg_only)
            self.clip_model, self.clip_layer, self.fuser = clip_model, clip_layer, None
            self.image_mean = torch.tensor(feature_extractor.image_mean, device=model.device).view(1, 3, 1, 1)
            self.image_std = torch.tensor(feature_extractor.image_std, device=model.device).view(1, 3, 1, 1)
            size = feature_extractor.size
            self.clip_size = size["shortest_edge"] if isinstance(size, dict) else size
            return
        channels = model.cin // 64               # (`model.config.cin // 64`, unified_pipeline.py:869)
        img, mask = _split_hint(image, mask, channels=1 if channels == 1 else 3)
        self.image = img.to(device=model.device, dtype=torch.float16).contiguous()
        self.mask = None if mask is None else mask.to(device=model.device, dtype=torch.float16).contiguous()
        self._resized = _MaskResizer()
        self.weight = float(weight)
        self.soft_injection = bool(soft_injection)
        self.cfg_only = bool(cfg_only)
        self.fuser = None

    def to(self, device=None, dtype=None):
        return self

    def coadapter_type(self):
        return False

    def style_call(self):
        """:941-975: rescale to the CLIP size ("cover"), normalise, vision tower hidden state of the chosen layer, adapter."""
        from .images import rescale
        layer = normalise_clip_layer(self.clip_layer, "final")
        image = rescale(self.image, self.clip_size, self.clip_size, "cover")
        image = (image - self.image_mean) / self.image_std
        out = self.clip_model.vision_model(image, output_hidden_states=(layer != "final"), return_dict=True)
        if layer == "final":
            hidden = out.last_hidden_state
        elif layer == "penultimate":
            hidden = out.hidden_states[-2]
        else:
            hidden = out.hidden_states[-layer]
        return self.model(hidden) * self.weight

    def __call__(self):
        if self.type == "style":
            return self.style_call()
        layer_weights = [1.0, 1.0, 1.0, 1.0]
        if self.soft_injection:
            layer_weights = torch.logspace(-0.25, 0, 4).tolist()
            if self.cfg_only:
                layer_weights[0] = 0.25
        states = self.model(self.image)
        if len(states) != 4:
            layer_weights = [1.0] * len(states)
        out = []
        for state, lw in zip(states, layer_weights):
            s = self.weight * lw
            state = state if s == 1.0 else state * s
            if self.mask is not None:
                state = state * self._resized(state, self.mask)
            out.append(state)
        return out


class AdapterStateList:
    """core.py:67-94: adapter states with their cfg_only flag; `all` feeds the guided side, `either` the unconditional one
    (cfg_only states become zeros there)."""

    def __init__(self):
        self.items = []

    def append(self, item, cfg_only: bool):
        self.items.append((item, cfg_only))

    @staticmethod
    def _zeros_like(item):
        return torch.zeros_like(item) if isinstance(item, torch.Tensor) else [AdapterStateList._zeros_like(s) for s in item]

    @property
    def all(self):
        return (item for item, _ in self.items)

    @property
    def cfg_only(self):
        return (item if cfg_only else self._zeros_like(item) for item, cfg_only in self.items)

    @property
    def either(self):
        return (item if not cfg_only else self._zeros_like(item) for item, cfg_only in self.items)


def _sum_lists(lists):
    return [sum(parts) for parts in zip(*lists)]


def controlnet_residual_kwargs(controlnets, latents, t, encoder_hidden_states, cfg_meta):
    """UNetWithControlnet.__call__ (core.py:44-64): every ControlNet sees the same latents / timestep / embeddings, their
    residuals are summed layer by layer."""
    res = [cn(latents, t, encoder_hidden_states=encoder_hidden_states, cfg_meta=cfg_meta) for cn in controlnets]
    return {"down_block_additional_residuals": _sum_lists([r.down_block_res_samples for r in res]),
            "mid_block_additional_residual": sum(r.mid_block_res_sample for r in res)}


class UNetWithControlnet:
    def __init__(self, unet, controlnets):
        self.unet = unet
        self.controlnets = controlnets

    def __call__(self, latents, t, **kwargs):
        resargs = controlnet_residual_kwargs(self.controlnets, latents, t, kwargs.get("encoder_hidden_states"),
                                             kwargs.get("cfg_meta"))
        return self.unet(latents, t, **kwargs, **resargs)


class UNetWithT2I:
    """core.py:97-239 for standard adapters: the states are computed ONCE (they depend on the hint image only), summed over
    adapters, and selected per call by `cfg_meta` ("u" / "g" / "f" = [u ; g])."""

    def __init__(self, unet, t2i_adapters):
        self.unet = unet
        self.standard_states = None
        self.style_states = None
        standard = AdapterStateList()
        style = []
        for adapter in t2i_adapters:
            if adapter.coadapter_type():
                raise NotImplementedError("co-adapters need the fuser model: not built")
            state = adapter()
            if isinstance(state, list):
                standard.append(state, adapter.cfg_only)
            else:
                style.append(state)              # style states always apply to the guided side only (core.py:99-100)
        if style:
            self.style_states = torch.cat(style, dim=1)
            self.style_dim0, self.style_dim1 = self.style_states.shape[0], self.style_states.shape[1]
        g, u = list(standard.all), list(standard.either)
        if g:
            self.standard_states = {"u": _sum_lists(u), "g": _sum_lists(g)}
            self.standard_dim0 = self.standard_states["g"][0].shape[0]
            self.standard_states["f"] = [torch.cat([a, b], dim=0) for a, b in zip(self.standard_states["u"], self.standard_states["g"])]

    def states_for(self, cfg_meta, batch):
        """The states for a UNet batch of `batch` rows: one hint image serves every sample of the request."""
        if self.standard_states is None:
            return None
        states = self.standard_states[cfg_meta]
        rows = batch // 2 if cfg_meta == "f" else batch
        if states[0].shape[0] == (2 if cfg_meta == "f" else 1) and rows > 1:
            if cfg_meta == "f":
                states = [torch.cat([s[:1].expand(rows, -1, -1, -1), s[1:].expand(rows, -1, -1, -1)]).contiguous() for s in states]
            else:
                states = [s.expand(rows, -1, -1, -1).contiguous() for s in states]
            self.standard_states = dict(self.standard_states, **{cfg_meta: states})      # expanded once per request
        return states

    def styled_context(self, hidden_states, cfg_meta):
        """core.py:221-237: the guided side's context gets the style tokens appended; the unconditional side is padded to
        the same length with its own last tokens."""
        if self.style_states is None:
            return hidden_states
        if cfg_meta == "f":
            uncond, cond = hidden_states.chunk(2)
        elif cfg_meta == "u":
            uncond, cond = hidden_states, None
        else:
            uncond, cond = None, hidden_states
        res = []
        if uncond is not None:
            res.append(torch.cat([uncond, uncond[:, -self.style_dim1:, :]], dim=1))
        if cond is not None:
            style = self.style_states.to(cond.dtype)
            if style.shape[0] == 1 and cond.shape[0] > 1:
                style = style.expand(cond.shape[0], -1, -1)          # one hint image serves every sample of the request
            res.append(torch.cat([cond, style], dim=1))
        return torch.cat(res, dim=0)

    def __call__(self, latents, t, **kwargs):
        dim0 = self.standard_dim0 if self.standard_states is not None else self.style_dim0
        is_f = kwargs["encoder_hidden_states"].shape[0] == dim0 * 2
        cfg_meta = kwargs.get("cfg_meta", "f" if is_f else "g")
        if self.standard_states is not None:
            kwargs["adapter_states"] = self.standard_states[cfg_meta]
        if self.style_states is not None:
            kwargs["encoder_hidden_states"] = self.styled_context(kwargs.pop("encoder_hidden_states"), cfg_meta)
        return self.unet(latents, t, **kwargs)
