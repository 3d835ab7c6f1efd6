"""Safety checker on the native kernels (the step right after the VAE decode).

Host-side mirror of what the reference calls at gyre/pipeline/unified_pipeline.py:2514-2522:

    safety_cheker_input = self.feature_extractor(self.numpy_to_pil(result_numpy), return_tensors="pt").to(device)
    result_numpy, has_nsfw_concept = self.safety_checker(images=result_numpy,
                                                         clip_input=safety_cheker_input.pixel_values.to(latents_dtype))

`B200FeatureExtractor` is the CLIPFeatureExtractor call (transformers ~= 4.28: shortest edge -> 224 with Pillow's BICUBIC,
centre crop, x 1/255, normalise) on the device: Pillow's 8-bit two-pass resample with the coefficient tables built here on
the host with Pillow's expressions (Resample.c precompute_coeffs / normalize_coeffs_8bpc) and applied by
gyre_b200_resample_u8 - bit-exact against `PIL.Image.resize`, so the checker sees the pixel values it was calibrated on -
and it takes the decoded image straight from HBM instead of a host round trip through PIL.

`B200SafetyChecker` is gyre/pipeline/safety_checkers.py:13-66 FlagOnlySafetyChecker (same state-dict names, same
`forward(clip_input, images) -> (images, has_nsfw_concepts)`): the CLIP ViT vision tower, visual projection and cosine
scores run in gyre_b200_safety_scores; the thresholding over the 20 scores per image is the reference's host loop."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _native as N

_PRECISION_BITS = 32 - 8 - 2          # Pillow: 8-bit pixels, 2 bits of headroom for the negative bicubic lobes
CLIP_IMAGE_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_IMAGE_STD = (0.26862954, 0.26130258, 0.27577711)


def pil_bicubic_tables(in_size: int, out_size: int):
    """Pillow's coefficient tables for one axis (Resample.c: precompute_coeffs with the bicubic filter, a = -0.5, support
    2; normalize_coeffs_8bpc): bounds [out, 2] int32 (first tap, tap count) and coeffs [out, ksize] int32, 22-bit fixed
    point.  All arithmetic in float64 in the order the C code uses, so the integers match Pillow's."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)              # (int) truncation of non-negative values
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    taps = np.arange(ksize, dtype=np.float64)[None, :]
    x = np.abs((taps + xmin[:, None] - center[:, None] + 0.5) * inv)
    a = -0.5
    w = np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1, np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))
    w = np.where(taps < xmax[:, None], w, 0.0)
    ww = np.zeros(out_size, np.float64)
    for j in range(ksize):                                                       # left-to-right double sum, like the C loop
        ww = ww + w[:, j]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    fixed = np.where(w < 0, np.trunc(-0.5 + w * (1 << _PRECISION_BITS)), np.trunc(0.5 + w * (1 << _PRECISION_BITS)))
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return bounds, fixed.astype(np.int32), ksize


class _Batch(dict):
    """BatchFeature's surface as the pipeline uses it: `.pixel_values`, `.to(device)`."""
    __getattr__ = dict.__getitem__

    def to(self, *args, **kwargs):
        return _Batch({k: v.to(*args, **kwargs) for k, v in self.items()})


class B200FeatureExtractor:
    """CLIPFeatureExtractor(do_resize, size=224 shortest edge, resample=BICUBIC, do_center_crop, crop_size=224,
    do_normalize) over device tensors."""

    def __init__(self, size: int = 224, crop_size: int | None = None, image_mean=CLIP_IMAGE_MEAN, image_std=CLIP_IMAGE_STD,
                 device=None):
        if isinstance(size, dict):
            size = size.get("shortest_edge", size.get("height"))
        if isinstance(crop_size, dict):
            crop_size = crop_size.get("height")
        self.size = int(size)
        self.crop_size = int(crop_size) if crop_size is not None else self.size
        if self.crop_size > self.size:
            raise ValueError("B200FeatureExtractor: crop_size larger than size would need padding (not built)")
        self.image_mean = tuple(float(v) for v in image_mean)
        self.image_std = tuple(float(v) for v in image_std)
        if not torch.cuda.is_available():
            raise N.NativeError("B200FeatureExtractor needs a CUDA device: there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = N.load()
        self._tables = {}

    def _axis_tables(self, in_size, out_size):
        key = (in_size, out_size)
        t = self._tables.get(key)
        if t is None:
            bounds, coeffs, ksize = pil_bicubic_tables(in_size, out_size)
            t = (torch.from_numpy(bounds).to(self.device), torch.from_numpy(coeffs).to(self.device), ksize)
            self._tables[key] = t
        return t

    def output_size(self, h, w):
        short, long = (w, h) if w <= h else (h, w)
        new_long = int(self.size * long / short)
        return (new_long, self.size) if w <= h else (self.size, new_long)

    def resize(self, u8_nhwc, out_h, out_w):
        """Pillow `Image.resize((out_w, out_h), BICUBIC)` of a batch of 8-bit RGB images [B, H, W, 3] on the device."""
        B, H, W, ch = u8_nhwc.shape
        st = N.stream_ptr(self.device)
        cur = u8_nhwc
        if out_w != W:
            bounds, coeffs, ks = self._axis_tables(W, out_w)
            dst = torch.empty((B, H, out_w, ch), device=self.device, dtype=torch.uint8)
            N.check(self._lib.gyre_b200_resample_u8(N.ptr(cur), B * H, W, ch, N.ptr(bounds), N.ptr(coeffs), ks, out_w,
                                                    N.ptr(dst), st), "resample_u8")
            cur = dst
        if out_h != H:
            bounds, coeffs, ks = self._axis_tables(H, out_h)
            dst = torch.empty((B, out_h, out_w, ch), device=self.device, dtype=torch.uint8)
            N.check(self._lib.gyre_b200_resample_u8(N.ptr(cur), B, H, out_w * ch, N.ptr(bounds), N.ptr(coeffs), ks, out_h,
                                                    N.ptr(dst), st), "resample_u8")
            cur = dst
        return cur

    def __call__(self, images, return_tensors="pt"):
        """images: float [B, 3, H, W] in [0, 1] (the pipeline's decoded images, quantised like numpy_to_pil:
        (x * 255).round()) or uint8 [B, H, W, 3].  Returns `.pixel_values` fp16 [B, 3, crop, crop] on the device."""
        if return_tensors != "pt":
            raise ValueError("B200FeatureExtractor returns torch tensors only")
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(images)
        if images.dtype != torch.uint8:
            from .images import to_uint8_nhwc
            images = to_uint8_nhwc(images.to(self.device))
        u8 = images.to(self.device).contiguous()
        if u8.ndim != 4 or u8.shape[-1] != 3:
            raise ValueError(f"B200FeatureExtractor: want [B, H, W, 3] uint8 or [B, 3, H, W] float, got {tuple(u8.shape)}")
        B, H, W, _ = u8.shape
        nh, nw = self.output_size(H, W)
        with torch.cuda.device(self.device):
            r = self.resize(u8, nh, nw)
            out = torch.empty((B, 3, self.crop_size, self.crop_size), device=self.device, dtype=torch.float16)
            mean = (C.c_float * 3)(*self.image_mean)
            std = (C.c_float * 3)(*self.image_std)
            N.check(self._lib.gyre_b200_clip_normalize(N.ptr(r), B, nh, nw, self.crop_size, mean, std, N.ptr(out),
                                                       N.stream_ptr(self.device)), "clip_normalize")
        return _Batch(pixel_values=out)


@dataclass
class ClipVisionConfig:
    image_size: int = 224
    patch_size: int = 14
    hidden_size: int = 1024
    intermediate_size: int = 4096
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    hidden_act: str = "quick_gelu"
    layer_norm_eps: float = 1e-5
    projection_dim: int = 768
    num_concepts: int = 17
    num_special: int = 3

    @staticmethod
    def vit_l14():
        """CompVis/stable-diffusion-safety-checker (CLIP ViT-L/14)."""
        return ClipVisionConfig()

    @staticmethod
    def tiny():
        return ClipVisionConfig(image_size=56, patch_size=14, hidden_size=64, intermediate_size=256, num_hidden_layers=2,
                                num_attention_heads=4, projection_dim=32)

    @classmethod
    def from_any(cls, cfg):
        """Accepts this class, a dict, or a transformers CLIPConfig (vision_config + projection_dim)."""
        if isinstance(cfg, cls):
            return cfg
        get = (lambda o, k: o.get(k)) if isinstance(cfg, dict) else (lambda o, k: getattr(o, k, None))
        vis = get(cfg, "vision_config")
        out = cls()
        src = vis if vis is not None else cfg
        vget = (lambda k: src.get(k)) if isinstance(src, dict) else (lambda k: getattr(src, k, None))
        for f in cls.__dataclass_fields__:
            v = vget(f)
            if v is not None:
                setattr(out, f, v)
        pd = get(cfg, "projection_dim")
        if pd is not None:
            out.projection_dim = pd
        return out


def safety_checker_param_shapes(cfg) -> dict:
    """FlagOnlySafetyChecker state-dict names -> shapes."""
    cfg = ClipVisionConfig.from_any(cfg)
    Cc, F, P = cfg.hidden_size, cfg.intermediate_size, cfg.patch_size
    ntok = (cfg.image_size // P) ** 2 + 1
    ks = {"vision_model.vision_model.embeddings.class_embedding": (Cc,),
          "vision_model.vision_model.embeddings.patch_embedding.weight": (Cc, 3, P, P),
          "vision_model.vision_model.embeddings.position_embedding.weight": (ntok, Cc),
          "vision_model.vision_model.pre_layrnorm.weight": (Cc,), "vision_model.vision_model.pre_layrnorm.bias": (Cc,),
          "vision_model.vision_model.post_layernorm.weight": (Cc,), "vision_model.vision_model.post_layernorm.bias": (Cc,),
          "visual_projection.weight": (cfg.projection_dim, Cc),
          "concept_embeds": (cfg.num_concepts, cfg.projection_dim), "special_care_embeds": (cfg.num_special, cfg.projection_dim),
          "concept_embeds_weights": (cfg.num_concepts,), "special_care_embeds_weights": (cfg.num_special,)}
    for i in range(cfg.num_hidden_layers):
        p = f"vision_model.vision_model.encoder.layers.{i}"
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


class B200SafetyChecker:
    """FlagOnlySafetyChecker (gyre/pipeline/safety_checkers.py:13-66)."""

    def __init__(self, config, device=None):
        self.config = ClipVisionConfig.from_any(config)
        if not torch.cuda.is_available():
            raise N.NativeError("B200SafetyChecker needs a CUDA device: there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = torch.float16
        self._lib = N.load()
        self._h = C.c_void_p()
        self._ws = {}
        self._loaded = False
        self._special_thresholds = None
        self._concept_thresholds = None
        self.last_result = None
        cfg = self.config
        if cfg.hidden_act not in ("quick_gelu", "gelu"):
            raise ValueError(f"hidden_act {cfg.hidden_act!r} not supported")
        c = N.ClipVisionConfigC(cfg.image_size, cfg.patch_size, cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers,
                                cfg.num_attention_heads, 0 if cfg.hidden_act == "quick_gelu" else 1, cfg.layer_norm_eps,
                                cfg.projection_dim, cfg.num_concepts, cfg.num_special)
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_clip_vision_create(C.byref(c), C.byref(self._h)), "clip_vision_create")

    def __str__(self):
        return "B200SafetyChecker"

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.gyre_b200_destroy(h)
            except Exception:
                pass
            self._h = None

    def load_state_dict(self, state_dict, strict: bool = True):
        """Keys as FlagOnlySafetyChecker / StableDiffusionSafetyChecker save them: the tower sits under
        `vision_model.vision_model.` (a CLIPVisionModel inside the checker); a bare CLIPVisionModel prefix works too."""
        for k, v in state_dict.items():
            if k.endswith("position_ids"):
                continue                      # buffer, not a parameter
            key = k[len("vision_model."):] if k.startswith("vision_model.vision_model.") else k
            t = v.detach()
            if key == "special_care_embeds_weights":
                self._special_thresholds = [float(x) for x in t.float().cpu()]
            elif key == "concept_embeds_weights":
                self._concept_thresholds = [float(x) for x in t.float().cpu()]
            if t.dtype not in (torch.float16, torch.float32):
                t = t.float()
            t = t.to(self.device).contiguous()
            shape = (C.c_int64 * t.ndim)(*t.shape)
            with torch.cuda.device(self.device):
                N.check(self._lib.gyre_b200_load_weight(self._h, key.encode(), N.ptr(t), N.dtype_code(t), shape, t.ndim,
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

    def scores(self, clip_input, return_embeds: bool = False):
        """[B, num_special + num_concepts] fp32 cosine scores (special-care columns first) on the device."""
        if not self._loaded:
            raise N.NativeError("B200SafetyChecker: weights not loaded")
        S = self.config.image_size
        x = clip_input.to(device=self.device, dtype=torch.float16).contiguous()
        if x.ndim != 4 or tuple(x.shape[1:]) != (3, S, S):
            raise ValueError(f"B200SafetyChecker: clip_input must be [B, 3, {S}, {S}], got {tuple(x.shape)}")
        B = x.shape[0]
        ne = self.config.num_special + self.config.num_concepts
        scores = torch.empty((B, ne), device=self.device, dtype=torch.float32)
        embeds = torch.empty((B, self.config.projection_dim), device=self.device, dtype=torch.float16)
        ws = self._workspace(B)
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_safety_scores(self._h, N.ptr(x), B, N.ptr(embeds), N.ptr(scores), N.ptr(ws), ws.numel(),
                                                      N.stream_ptr(self.device)), "safety_scores")
        return (scores, embeds) if return_embeds else scores

    def forward(self, clip_input, images):
        if self._special_thresholds is None or self._concept_thresholds is None:
            raise N.NativeError("B200SafetyChecker: concept thresholds not loaded")
        cos = self.scores(clip_input).cpu().numpy()           # B x 20 floats: the only device -> host read of this stage
        ns = self.config.num_special
        result = []
        for row in cos:
            res = {"special_scores": {}, "special_care": [], "concept_scores": {}, "bad_concepts": []}
            adjustment = 0.0
            for i, thr in enumerate(self._special_thresholds):
                res["special_scores"][i] = round(row[i] - thr + adjustment, 3)
                if res["special_scores"][i] > 0:
                    res["special_care"].append({i, res["special_scores"][i]})
                    adjustment = 0.01
            for i, thr in enumerate(self._concept_thresholds):
                res["concept_scores"][i] = round(row[ns + i] - thr + adjustment, 3)
                if res["concept_scores"][i] > 0:
                    res["bad_concepts"].append(i)
            result.append(res)
        self.last_result = result
        return images, [len(r["bad_concepts"]) > 0 for r in result]

    __call__ = forward
