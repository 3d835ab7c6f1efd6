"""CLIP text encoder on the native kernels (the step right before the sampling path).

Host-side mirror of what the reference calls: `CLIPTextModel(input_ids, output_hidden_states=, return_dict=True)`
(transformers, third-party) through `TextEncoderAltLayer` (gyre/pipeline/text_embedding/text_encoder_alt_layer.py:6-36),
which picks `last_hidden_state` ("final"), `final_layer_norm(hidden_states[-2])` ("penultimate", SD2.x) or
`final_layer_norm(hidden_states[-layer])`.  Tokenisation / prompt weighting (lpw_text_embedding.py) stay in Python
upstream of this class: it takes token ids."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _native as N


@dataclass
class ClipTextConfig:
    vocab_size: int = 49408
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    max_position_embeddings: int = 77
    hidden_act: str = "quick_gelu"
    layer_norm_eps: float = 1e-5

    @staticmethod
    def clip_l():
        return ClipTextConfig()

    @staticmethod
    def open_clip_h():
        """SD2.x text encoder (OpenCLIP ViT-H/14 text tower as shipped in diffusers format: 23 layers kept)."""
        return ClipTextConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=23, num_attention_heads=16,
                              hidden_act="gelu")

    @staticmethod
    def tiny():
        return ClipTextConfig(vocab_size=1000, hidden_size=64, intermediate_size=256, num_hidden_layers=3,
                              num_attention_heads=4)

    @classmethod
    def from_any(cls, cfg):
        if isinstance(cfg, cls):
            return cfg
        get = (lambda k: cfg.get(k)) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k, None))
        out = cls()
        for f in cls.__dataclass_fields__:
            v = get(f)
            if v is not None:
                setattr(out, f, v)
        return out


def clip_param_shapes(cfg) -> dict:
    """transformers CLIPTextModel state-dict names -> shapes."""
    cfg = ClipTextConfig.from_any(cfg)
    Cc, F = cfg.hidden_size, cfg.intermediate_size
    ks = {"text_model.embeddings.token_embedding.weight": (cfg.vocab_size, Cc),
          "text_model.embeddings.position_embedding.weight": (cfg.max_position_embeddings, Cc),
          "text_model.final_layer_norm.weight": (Cc,), "text_model.final_layer_norm.bias": (Cc,)}
    for i in range(cfg.num_hidden_layers):
        p = f"text_model.encoder.layers.{i}"
        for nme in ("q_proj", "k_proj", "v_proj", "out_proj"):
            ks[f"{p}.self_attn.{nme}.weight"] = (Cc, Cc)
            ks[f"{p}.self_attn.{nme}.bias"] = (Cc,)
        for nme in ("layer_norm1", "layer_norm2"):
            ks[f"{p}.{nme}.weight"] = (Cc,)
            ks[f"{p}.{nme}.bias"] = (Cc,)
        ks[f"{p}.mlp.fc1.weight"] = (F, Cc)
        ks[f"{p}.mlp.fc1.bias"] = (F,)
        ks[f"{p}.mlp.fc2.weight"] = (Cc, F)
        ks[f"{p}.mlp.fc2.bias"] = (Cc,)
    return ks


@dataclass
class TextEncoderOutput:
    last_hidden_state: torch.Tensor
    hidden_states: tuple | None = None


class B200CLIPTextModel:
    def __init__(self, config, device=None):
        self.config = ClipTextConfig.from_any(config)
        if not torch.cuda.is_available():
            raise N.NativeError("B200CLIPTextModel needs a CUDA device: there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = torch.float16
        self._lib = N.load()
        self._h = C.c_void_p()
        self._ws = {}
        self._loaded = False
        cfg = self.config
        if cfg.hidden_act not in ("quick_gelu", "gelu"):
            raise ValueError(f"hidden_act {cfg.hidden_act!r} not supported")
        c = N.ClipConfigC(cfg.vocab_size, cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers,
                          cfg.num_attention_heads, cfg.max_position_embeddings, 0 if cfg.hidden_act == "quick_gelu" else 1,
                          cfg.layer_norm_eps)
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_clip_create(C.byref(c), C.byref(self._h)), "clip_create")

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
            if k.endswith("position_ids"):
                continue                      # buffer, not a parameter
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

    def _workspace(self, B, L):
        ws = self._ws.get((B, L))
        if ws is None:
            n = C.c_size_t()
            N.check(self._lib.gyre_b200_clip_workspace_bytes(self._h, B, L, C.byref(n)), "clip_workspace_bytes")
            self._ws.clear()
            ws = torch.empty((n.value,), device=self.device, dtype=torch.uint8)
            self._ws[(B, L)] = ws
        return ws

    def encode(self, input_ids, layer="final", apply_final_ln=True):
        """TextEncoderAltLayer semantics: "final" | "penultimate" | int n -> final_layer_norm(hidden_states[-n])."""
        if not self._loaded:
            raise N.NativeError("B200CLIPTextModel: weights not loaded")
        skip = 0 if layer == "final" else (1 if layer == "penultimate" else int(layer) - 1)
        ids = input_ids.to(device=self.device, dtype=torch.int64).contiguous()
        B, L = ids.shape
        out = torch.empty((B, L, self.config.hidden_size), device=self.device, dtype=torch.float16)
        ws = self._workspace(B, L)
        N.check(self._lib.gyre_b200_clip_forward(self._h, N.ptr(ids), B, L, skip, 1 if apply_final_ln else 0, N.ptr(out),
                                                 N.ptr(ws), ws.numel(), N.stream_ptr(self.device)), "clip_forward")
        return out

    def __call__(self, input_ids, output_hidden_states=False, return_dict=True, **_):
        last = self.encode(input_ids, "final")
        hs = None
        if output_hidden_states:
            # hidden_states[k] = output of k layers (no final LN), k = 0 .. num_layers (transformers convention)
            nl = self.config.num_hidden_layers
            hs = tuple(self.encode(input_ids, layer=nl - k + 1, apply_final_ln=False) for k in range(nl + 1))
        return TextEncoderOutput(last_hidden_state=last, hidden_states=hs)
