"""Oracle: the safety-checker tail of the pipeline restated on the CPU (TEST ONLY - nothing under gyre_b200/ imports it).

Three pieces, each following the code the reference runs:
  * `clip_preprocess`   - the CLIPFeatureExtractor call at gyre/pipeline/unified_pipeline.py:2516-2518 (transformers
    ~= 4.28.1, /root/reference/pyproject.toml:21: resize shortest edge -> 224 with PIL BICUBIC, centre crop 224, x 1/255,
    (x - mean) / std).  The resize is Pillow's 8-bit two-pass resample (src/libImaging/Resample.c: precompute_coeffs,
    normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc) restated in numpy.  PINNED against
    `PIL.Image.resize` itself (installed Pillow) by scripts/make_golden.py:pin_safety - bit-exact.
  * `clip_vision_forward` - transformers `CLIPVisionModel` (CLIPVisionTransformer) + the checker's visual projection,
    gyre/pipeline/safety_checkers.py:32-33.  PINNED against the reference class FlagOnlySafetyChecker built on the
    installed transformers (5.5; the vision tower's arithmetic is unchanged since 4.28).
  * `flag_only`         - FlagOnlySafetyChecker.forward's scoring loop, safety_checkers.py:35-66.  PINNED the same way."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def synthetic_image(h: int, w: int) -> np.ndarray:
    """A reproducible 8-bit RGB test image from integer arithmetic only (no RNG, no libm): smooth ramps, a hashed noise
    field and a block of hard 0/255 edges, so the interpolation, the negative bicubic lobes and the clamp all matter."""
    y, x = np.mgrid[0:h, 0:w].astype(np.int64)
    img = np.empty((h, w, 3), np.uint8)
    for c in range(3):
        ramp = (x * (3 + c) + y * (5 - c)) % 512
        ramp = np.where(ramp > 255, 511 - ramp, ramp)                              # triangle wave 0..255
        hsh = (x * 73856093) ^ (y * 19349663) ^ ((c + 1) * 83492791)
        noise = ((hsh >> 7) % 61) - 30
        img[..., c] = np.clip(ramp + noise, 0, 255)
    blk = (((x // 3) + (y // 2)) % 2 * 255).astype(np.uint8)
    img[: h // 4, : w // 4] = blk[: h // 4, : w // 4, None]
    return img


def _bicubic(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter (support 2) over the whole axis.
    Returns (bounds [out, 2] int32 = (first source index, tap count), coeffs [out, ksize] int32, ksize)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)                                   # C adds left to right in double: so does Python's sum over floats
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _resample_axis(img: np.ndarray, axis: int, out_size: int) -> np.ndarray:
    bounds, kk, _ = pil_bicubic_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for o in range(out_size):
        x0, n = bounds[o]
        k = kk[o, :n].astype(np.int64).reshape((n,) + (1,) * (src.ndim - 1))
        acc = (1 << (PRECISION_BITS - 1)) + (src[x0:x0 + n] * k).sum(0)
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bicubic(img_hwc_u8: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Image.resize((out_w, out_h), BICUBIC) for an 8-bit RGB image: horizontal pass, then vertical pass, each skipped
    when that size is unchanged (ImagingResample)."""
    out = img_hwc_u8
    if out_w != out.shape[1]:
        out = _resample_axis(out, 1, out_w)
    if out_h != out.shape[0]:
        out = _resample_axis(out, 0, out_h)
    return out


def resize_output_size(h: int, w: int, size: int = 224):
    """transformers get_resize_output_image_size(default_to_square=False): shortest edge -> size, the other int()-truncated."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)          # (new_h, new_w)


def clip_preprocess(images_u8_nhwc: np.ndarray, size: int = 224, mean=CLIP_MEAN, std=CLIP_STD) -> np.ndarray:
    """u8 [B, H, W, 3] -> float32 [B, 3, size, size] pixel_values (CLIPImageProcessor.preprocess, 4.28)."""
    out = []
    mean = np.array(mean, dtype=np.float32)
    std = np.array(std, dtype=np.float32)
    for img in images_u8_nhwc:
        nh, nw = resize_output_size(img.shape[0], img.shape[1], size)
        r = pil_resize_bicubic(img, nw, nh)
        top, left = (nh - size) // 2, (nw - size) // 2
        r = r[top:top + size, left:left + size]
        x = (r * (1 / 255)).astype(np.float32)                                   # u8 * python float -> float64 -> float32
        x = (x - mean) / std
        out.append(x.transpose(2, 0, 1))
    return np.stack(out)


def clip_vision_forward(P: dict, pixel_values, *, num_layers, num_heads, patch_size, hidden_act="quick_gelu", eps=1e-5,
                        return_hidden_states=False):
    """Returns (pooled_output [B, C], image_embeds [B, projection_dim]); with return_hidden_states also the transformers
    `hidden_states` tuple (entry k = output of k encoder layers, before post_layernorm; the last one is last_hidden_state)."""
    w = P["vision_model.embeddings.patch_embedding.weight"]
    x = F.conv2d(pixel_values, w, stride=patch_size).flatten(2).transpose(1, 2)             # [B, np, C]
    B, _, C = x.shape
    cls = P["vision_model.embeddings.class_embedding"].expand(B, 1, C)
    h = torch.cat([cls, x], dim=1) + P["vision_model.embeddings.position_embedding.weight"][None]
    h = F.layer_norm(h, (C,), P["vision_model.pre_layrnorm.weight"], P["vision_model.pre_layrnorm.bias"], eps)
    N = h.shape[1]
    d = C // num_heads
    hs = [h]
    for i in range(num_layers):
        p = f"vision_model.encoder.layers.{i}"
        n = F.layer_norm(h, (C,), P[f"{p}.layer_norm1.weight"], P[f"{p}.layer_norm1.bias"], eps)
        q = F.linear(n, P[f"{p}.self_attn.q_proj.weight"], P[f"{p}.self_attn.q_proj.bias"]) * d ** -0.5
        k = F.linear(n, P[f"{p}.self_attn.k_proj.weight"], P[f"{p}.self_attn.k_proj.bias"])
        v = F.linear(n, P[f"{p}.self_attn.v_proj.weight"], P[f"{p}.self_attn.v_proj.bias"])
        sp = lambda t: t.reshape(B, N, num_heads, d).permute(0, 2, 1, 3)
        s = sp(q) @ sp(k).transpose(-1, -2)
        o = (torch.softmax(s, dim=-1) @ sp(v)).permute(0, 2, 1, 3).reshape(B, N, C)
        h = h + F.linear(o, P[f"{p}.self_attn.out_proj.weight"], P[f"{p}.self_attn.out_proj.bias"])
        n = F.layer_norm(h, (C,), P[f"{p}.layer_norm2.weight"], P[f"{p}.layer_norm2.bias"], eps)
        m = F.linear(n, P[f"{p}.mlp.fc1.weight"], P[f"{p}.mlp.fc1.bias"])
        m = m * torch.sigmoid(1.702 * m) if hidden_act == "quick_gelu" else F.gelu(m)
        h = h + F.linear(m, P[f"{p}.mlp.fc2.weight"], P[f"{p}.mlp.fc2.bias"])
        hs.append(h)
    pooled = F.layer_norm(h[:, 0], (C,), P["vision_model.post_layernorm.weight"], P["vision_model.post_layernorm.bias"], eps)
    emb = F.linear(pooled, P["visual_projection.weight"]) if "visual_projection.weight" in P else None
    return (pooled, emb, hs) if return_hidden_states else (pooled, emb)


def cosine_scores(image_embeds, P):
    """[B, n_special + n_concepts]: special-care columns first (safety_checkers.py:8-11, 35-36)."""
    e = F.normalize(image_embeds.float())
    return torch.cat([e @ F.normalize(P["special_care_embeds"].float()).t(), e @ F.normalize(P["concept_embeds"].float()).t()], dim=1)


def flag_only(scores: np.ndarray, special_thresholds, concept_thresholds):
    """safety_checkers.py:39-66 on the cosine scores: per image the rounded margins and the nsfw flag."""
    ns = len(special_thresholds)
    result = []
    for row in scores:
        adjustment = 0.0
        r = {"special_scores": {}, "special_care": [], "concept_scores": {}, "bad_concepts": []}
        for i in range(ns):
            r["special_scores"][i] = round(row[i] - float(special_thresholds[i]) + adjustment, 3)
            if r["special_scores"][i] > 0:
                r["special_care"].append({i, r["special_scores"][i]})
                adjustment = 0.01
        for i in range(len(concept_thresholds)):
            r["concept_scores"][i] = round(row[ns + i] - float(concept_thresholds[i]) + adjustment, 3)
            if r["concept_scores"][i] > 0:
                r["bad_concepts"].append(i)
        result.append(r)
    return result, [len(r["bad_concepts"]) > 0 for r in result]
