"""CLIP vision tower and the style T2I-adapter on the native kernels (style hints).

Mirrors what `UnifiedPipelineHint_T2i.style_call` (gyre/pipeline/unified_pipeline.py:941-975) touches:
`clip_model.vision_model(image, output_hidden_states=, return_dict=True)` -> `.last_hidden_state` / `.hidden_states[-k]`
(transformers CLIPVisionTransformer; hidden states are taken BEFORE post_layernorm) and `StyleAdapter.forward`
(gyre/pipeline/t2i_adapter/adapter.py:173-199; `T2iAdapter_style`, models.py:146-160)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _native as N
from .safety_checker import ClipVisionConfig


class _LazyHiddenStates:
    """`hidden_states` of the transformers output: indexable, each entry computed on demand (the caller takes one)."""

    def __init__(self, model, pixel_values):
        self._m, self._x = model, pixel_values

    def __len__(self):
        return self._m.config.num_hidden_layers + 1

    def __getitem__(self, k):
        n = len(self)
        k = k + n if k < 0 else k
        if not 0 <= k < n:
            raise IndexError(k)
        return self._m.hidden(self._x, skip_last=n - 1 - k)


@dataclass
class VisionOutput:
    last_hidden_state: torch.Tensor
    hidden_states: object = None


class B200CLIPVisionModel:
    """`CLIPModel.vision_model` / `CLIPVisionModel`: state-dict keys `vision_model.*` (keys of a full CLIPModel that do not
    belong to the tower are skipped)."""

    def __init__(self, config, device=None):
        cfg = ClipVisionConfig.from_any(config)
        cfg.projection_dim, cfg.num_concepts, cfg.num_special = 0, 0, 0
        self.config = cfg
        if not torch.cuda.is_available():
            raise N.NativeError("B200CLIPVisionModel needs a CUDA device: there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = torch.float16
        self._lib = N.load()
        self._h = C.c_void_p()
        self._ws = {}
        self._loaded = False
        if cfg.hidden_act not in ("quick_gelu", "gelu"):
            raise ValueError(f"hidden_act {cfg.hidden_act!r} not supported")
        c = N.ClipVisionConfigC(cfg.image_size, cfg.patch_size, cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers,
                                cfg.num_attention_heads, 0 if cfg.hidden_act == "quick_gelu" else 1, cfg.layer_norm_eps, 0, 0, 0)
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_clip_vision_create(C.byref(c), C.byref(self._h)), "clip_vision_create")

    @property
    def vision_model(self):
        return self

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.gyre_b200_destroy(h)
            except Exception:
                pass
            self._h = None

    def load_state_dict(self, state_dict, strict: bool = True):
        for k, v in state_dict.items():
            if not k.startswith("vision_model.") or k.endswith("position_ids"):
                continue
            t = v.detach()
            if t.dtype not in (torch.float16, torch.float32):
                t = t.float()
            t = t.to(self.device).contiguous()
            shape = (C.c_int64 * t.ndim)(*t.shape)
            with torch.cuda.device(self.device):
                N.check(self._lib.gyre_b200_load_weight(self._h, k.encode(), N.ptr(t), N.dtype_code(t), shape, t.ndim,
                                                        N.stream_ptr(self.device)), f"load_weight({k})")
                torch.cuda.current_stream(self.device).synchronize()
        if strict:
            N.check(self._lib.gyre_b200_finalize(self._h), "finalize")
        self._loaded = True
        return self

    def _workspace(self, B):
        ws = self._ws.get(B)
        if ws is None:
            n = C.c_size_t()
            N.check(self._lib.gyre_b200_clip_vision_workspace_bytes(self._h, B, C.byref(n)), "clip_vision_workspace_bytes")
            self._ws.clear()
            ws = torch.empty((n.value,), device=self.device, dtype=torch.uint8)
            self._ws[B] = ws
        return ws

    def hidden(self, pixel_values, skip_last: int = 0):
        """Hidden states [B, tokens, C] after num_layers - skip_last encoder layers (no post_layernorm)."""
        if not self._loaded:
            raise N.NativeError("B200CLIPVisionModel: weights not loaded")
        S = self.config.image_size
        x = pixel_values.to(device=self.device, dtype=torch.float16).contiguous()
        if x.ndim != 4 or tuple(x.shape[1:]) != (3, S, S):
            raise ValueError(f"B200CLIPVisionModel: pixel_values must be [B, 3, {S}, {S}], got {tuple(x.shape)}")
        B = x.shape[0]
        ntok = (S // self.config.patch_size) ** 2 + 1
        out = torch.empty((B, ntok, self.config.hidden_size), device=self.device, dtype=torch.float16)
        ws = self._workspace(B)
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_clip_vision_hidden(self._h, N.ptr(x), B, int(skip_last), N.ptr(out), N.ptr(ws), ws.numel(),
                                                           N.stream_ptr(self.device)), "clip_vision_hidden")
        return out

    def __call__(self, pixel_values, output_hidden_states=False, return_dict=True, **_):
        return VisionOutput(last_hidden_state=self.hidden(pixel_values, 0),
                            hidden_states=_LazyHiddenStates(self, pixel_values) if output_hidden_states else None)


def style_adapter_param_shapes(width=1024, context_dim=768, num_head=8, n_layes=3, num_token=8) -> dict:
    ks = {"style_embedding": (1, num_token, width), "proj": (width, context_dim)}
    for n in ("ln_pre", "ln_post"):
        ks[f"{n}.weight"], ks[f"{n}.bias"] = (width,), (width,)
    for i in range(n_layes):
        p = f"transformer_layes.{i}"
        ks[f"{p}.attn.in_proj_weight"], ks[f"{p}.attn.in_proj_bias"] = (3 * width, width), (3 * width,)
        ks[f"{p}.attn.out_proj.weight"], ks[f"{p}.attn.out_proj.bias"] = (width, width), (width,)
        for n in ("ln_1", "ln_2"):
            ks[f"{p}.{n}.weight"], ks[f"{p}.{n}.bias"] = (width,), (width,)
        ks[f"{p}.mlp.c_fc.weight"], ks[f"{p}.mlp.c_fc.bias"] = (4 * width, width), (4 * width,)
        ks[f"{p}.mlp.c_proj.weight"], ks[f"{p}.mlp.c_proj.bias"] = (width, 4 * width), (width,)
    return ks


class B200T2iStyleAdapter:
    """`T2iAdapter_style(width, context_dim, num_head, n_layes, num_token)` (models.py:146-160; defaults of `type: style`:
    1024 / 768 / 8 / 3 / 8): CLIP vision hidden states [B, L, width] -> [B, num_token, context_dim] context tokens."""

    def __init__(self, width=1024, context_dim=768, num_head=8, n_layes=3, num_token=8, device=None):
        if not torch.cuda.is_available():
            raise N.NativeError("B200T2iStyleAdapter needs a CUDA device: there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = torch.float16
        self.width, self.context_dim, self.num_head, self.n_layes, self.num_token = width, context_dim, num_head, n_layes, num_token
        self._lib = N.load()
        self._h = C.c_void_p()
        self._ws = {}
        self._loaded = False
        c = N.StyleAdapterConfigC(width, context_dim, num_head, n_layes, num_token)
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_style_adapter_create(C.byref(c), C.byref(self._h)), "style_adapter_create")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.gyre_b200_destroy(h)
            except Exception:
                pass
            self._h = None

    def load_state_dict(self, state_dict, strict: bool = True):
        for k, v in state_dict.items():
            t = v.detach()
            if t.dtype not in (torch.float16, torch.float32):
                t = t.float()
            if k == "style_embedding":
                t = t.reshape(self.num_token, self.width)
            elif k == "proj":
                t = t.t()                       # the module multiplies x @ proj: the GEMM wants [context_dim, width]
            t = t.to(self.device).contiguous()
            shape = (C.c_int64 * t.ndim)(*t.shape)
            with torch.cuda.device(self.device):
                N.check(self._lib.gyre_b200_load_weight(self._h, k.encode(), N.ptr(t), N.dtype_code(t), shape, t.ndim,
                                                        N.stream_ptr(self.device)), f"load_weight({k})")
                torch.cuda.current_stream(self.device).synchronize()
        if strict:
            N.check(self._lib.gyre_b200_finalize(self._h), "finalize")
        self._loaded = True
        return self

    @torch.no_grad()
    def __call__(self, x):
        if not self._loaded:
            raise N.NativeError("B200T2iStyleAdapter: weights not loaded")
        N.require_cuda(x)
        if x.ndim != 3 or x.shape[2] != self.width:
            raise ValueError(f"style adapter input must be [B, tokens, {self.width}], got {tuple(x.shape)}")
        B, L, _ = x.shape
        x = x.to(torch.float16).contiguous()
        ws = self._ws.get((B, L))
        if ws is None:
            n = C.c_size_t()
            N.check(self._lib.gyre_b200_style_adapter_workspace_bytes(self._h, B, L, C.byref(n)), "style_adapter_workspace_bytes")
            self._ws.clear()
            ws = torch.empty((n.value,), device=self.device, dtype=torch.uint8)
            self._ws[(B, L)] = ws
        out = torch.empty((B, self.num_token, self.context_dim), device=self.device, dtype=torch.float16)
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_style_adapter_forward(self._h, N.ptr(x), B, L, N.ptr(out), N.ptr(ws), ws.numel(),
                                                              N.stream_ptr(self.device)), "style_adapter_forward")
        return out

    forward = __call__
