"""Oracle: diffusers-0.16 ``UNet2DConditionModel`` forward restated functionally (TEST ONLY).

The arithmetic lives in the third-party dependency ``diffusers ~= 0.16.0``
(/root/reference/pyproject.toml:22), which is not vendored; the in-tree statements this
follows are cited per function.  Weights are a flat ``dict[str, Tensor]`` whose keys are the
diffusers state-dict names, so a real checkpoint drops in unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from .tome import bipartite_soft_matching_plan, merge_mean, parse_r


@dataclass
class UNetConfig:
    """Hyper-parameters; SD1.x values follow gyre/ldm_config/v1-inference.yaml:29-44 and
    gyre/pipeline/controlnet/models.py:100-131 (defaults mirrored there)."""

    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: tuple = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    # diffusers-0.16 quirk: `attention_head_dim` is the NUMBER OF HEADS (per level)
    num_heads: tuple = (8, 8, 8, 8)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    attn_levels: tuple = (True, True, True, False)  # CrossAttnDownBlock2D x3, DownBlock2D
    sample_size: int = 64
    prediction_type: str = "epsilon"
    upcast_attention: bool = False
    # SDXL-style topologies (BASELINE config 5; public diffusers config names, not in the reference)
    transformer_layers_per_block: tuple = (1, 1, 1, 1)
    addition_time_embed_dim: int = 0
    projection_class_embeddings_input_dim: int = 0

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4

    @staticmethod
    def sd15(**kw):
        return UNetConfig(**kw)

    @staticmethod
    def sd15_inpaint(**kw):
        return UNetConfig(in_channels=9, **kw)

    @staticmethod
    def sd21_v(**kw):
        # gyre/ldm_config/v2-inference-v.yaml: num_head_channels 64, context_dim 1024, linear proj, v-pred
        return UNetConfig(num_heads=(5, 10, 20, 20), cross_attention_dim=1024, use_linear_projection=True,
                          sample_size=96, prediction_type="v_prediction", upcast_attention=True, **kw)

    @staticmethod
    def sdxl(**kw):
        return UNetConfig(block_out_channels=(320, 640, 1280), num_heads=(5, 10, 20), attn_levels=(False, True, True),
                          transformer_layers_per_block=(1, 2, 10), cross_attention_dim=2048, use_linear_projection=True,
                          sample_size=128, addition_time_embed_dim=256, projection_class_embeddings_input_dim=2816, **kw)

    @staticmethod
    def tiny_xl(**kw):
        """SDXL topology in miniature: 3 levels, no attention at level 0, transformer depth (1, 2, 3), head dim 32,
        text_time conditioning (pooled 32 + 6 x 16 sinusoids = 128)."""
        d = dict(block_out_channels=(64, 128, 256), num_heads=(2, 4, 8), attn_levels=(False, True, True),
                 transformer_layers_per_block=(1, 2, 3), cross_attention_dim=64, use_linear_projection=True,
                 sample_size=16, addition_time_embed_dim=16, projection_class_embeddings_input_dim=128)
        d.update(kw)
        return UNetConfig(**d)

    @staticmethod
    def tiny(**kw):
        """Same topology, 1/5 the width: CPU tests + committed golden fixtures."""
        d = dict(block_out_channels=(64, 128, 256, 256), num_heads=(4, 4, 4, 4), cross_attention_dim=64,
                 sample_size=16)
        d.update(kw)
        return UNetConfig(**d)


# ----------------------------------------------------------------------------- weights

def _resnet_keys(p, cin, cout, temb):
    ks = {
        f"{p}.norm1.weight": (cin,), f"{p}.norm1.bias": (cin,),
        f"{p}.conv1.weight": (cout, cin, 3, 3), f"{p}.conv1.bias": (cout,),
        f"{p}.norm2.weight": (cout,), f"{p}.norm2.bias": (cout,),
        f"{p}.conv2.weight": (cout, cout, 3, 3), f"{p}.conv2.bias": (cout,),
    }
    if temb:
        ks[f"{p}.time_emb_proj.weight"] = (cout, temb)
        ks[f"{p}.time_emb_proj.bias"] = (cout,)
    if cin != cout:
        ks[f"{p}.conv_shortcut.weight"] = (cout, cin, 1, 1)
        ks[f"{p}.conv_shortcut.bias"] = (cout,)
    return ks


def _transformer_keys(p, c, ctx, linear, depth=1):
    ks = {f"{p}.norm.weight": (c,), f"{p}.norm.bias": (c,)}
    if linear:
        ks[f"{p}.proj_in.weight"] = (c, c)
        ks[f"{p}.proj_out.weight"] = (c, c)
    else:
        ks[f"{p}.proj_in.weight"] = (c, c, 1, 1)
        ks[f"{p}.proj_out.weight"] = (c, c, 1, 1)
    ks[f"{p}.proj_in.bias"] = (c,)
    ks[f"{p}.proj_out.bias"] = (c,)
    for bi in range(depth):
        b = f"{p}.transformer_blocks.{bi}"
        for n in ("norm1", "norm2", "norm3"):
            ks[f"{b}.{n}.weight"] = (c,)
            ks[f"{b}.{n}.bias"] = (c,)
        for a, kd in (("attn1", c), ("attn2", ctx)):
            ks[f"{b}.{a}.to_q.weight"] = (c, c)
            ks[f"{b}.{a}.to_k.weight"] = (c, kd)
            ks[f"{b}.{a}.to_v.weight"] = (c, kd)
            ks[f"{b}.{a}.to_out.0.weight"] = (c, c)
            ks[f"{b}.{a}.to_out.0.bias"] = (c,)
        ks[f"{b}.ff.net.0.proj.weight"] = (8 * c, c)
        ks[f"{b}.ff.net.0.proj.bias"] = (8 * c,)
        ks[f"{b}.ff.net.2.weight"] = (c, 4 * c)
        ks[f"{b}.ff.net.2.bias"] = (c,)
    return ks


def _depth(cfg, level):
    d = getattr(cfg, "transformer_layers_per_block", None)
    return int(d[level]) if d else 1


def unet_param_shapes(cfg: UNetConfig) -> dict:
    """Every parameter name -> shape, in diffusers state-dict naming."""
    ch = cfg.block_out_channels
    T = cfg.time_embed_dim
    ks = {
        "conv_in.weight": (ch[0], cfg.in_channels, 3, 3), "conv_in.bias": (ch[0],),
        "time_embedding.linear_1.weight": (T, ch[0]), "time_embedding.linear_1.bias": (T,),
        "time_embedding.linear_2.weight": (T, T), "time_embedding.linear_2.bias": (T,),
    }
    if getattr(cfg, "addition_time_embed_dim", 0):
        pin = cfg.projection_class_embeddings_input_dim
        ks.update({"add_embedding.linear_1.weight": (T, pin), "add_embedding.linear_1.bias": (T,),
                   "add_embedding.linear_2.weight": (T, T), "add_embedding.linear_2.bias": (T,)})
    skips = [ch[0]]
    cin = ch[0]
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block):
            ks.update(_resnet_keys(f"down_blocks.{i}.resnets.{j}", cin, c, T))
            cin = c
            if cfg.attn_levels[i]:
                ks.update(_transformer_keys(f"down_blocks.{i}.attentions.{j}", c, cfg.cross_attention_dim,
                                            cfg.use_linear_projection, _depth(cfg, i)))
            skips.append(c)
        if i < len(ch) - 1:
            ks[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (c, c, 3, 3)
            ks[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (c,)
            skips.append(c)
    ks.update(_resnet_keys("mid_block.resnets.0", cin, cin, T))
    ks.update(_transformer_keys("mid_block.attentions.0", cin, cfg.cross_attention_dim, cfg.use_linear_projection,
                                _depth(cfg, len(ch) - 1)))
    ks.update(_resnet_keys("mid_block.resnets.1", cin, cin, T))
    rch = list(reversed(ch))
    rattn = list(reversed(cfg.attn_levels))
    for i, c in enumerate(rch):
        for j in range(cfg.layers_per_block + 1):
            s = skips.pop()
            ks.update(_resnet_keys(f"up_blocks.{i}.resnets.{j}", cin + s, c, T))
            cin = c
            if rattn[i]:
                ks.update(_transformer_keys(f"up_blocks.{i}.attentions.{j}", c, cfg.cross_attention_dim,
                                            cfg.use_linear_projection, _depth(cfg, len(ch) - 1 - i)))
        if i < len(ch) - 1:
            ks[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (c, c, 3, 3)
            ks[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (c,)
    ks["conv_norm_out.weight"] = (ch[0],)
    ks["conv_norm_out.bias"] = (ch[0],)
    ks["conv_out.weight"] = (cfg.out_channels, ch[0], 3, 3)
    ks["conv_out.bias"] = (cfg.out_channels,)
    return ks


# Residual-branch output layers get a reduced gain so that 50 chained forwards of a random-init
# network stay O(1) (the reference ships trained weights; synthetic ones must be well-conditioned).
_BRANCH_OUT = ("conv2.weight", "to_out.0.weight", "ff.net.2.weight", "proj_out.weight", "proj_attn.weight")


def synth_params(shapes: dict, seed: int, dtype=torch.float32, branch_gain: float = 0.5) -> dict:
    """Deterministic synthetic weights: W ~ N(0, gain^2/fan_in), norm gamma ~ 1 + 0.1 N, small biases.
    One CPU Philox generator walked in sorted-key order => identical on every box of this image."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for k in sorted(shapes):
        shp = shapes[k]
        if k.endswith(".weight") and len(shp) == 1:      # norm gamma
            w = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".bias"):
            w = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            gain = branch_gain if k.endswith(_BRANCH_OUT) else 1.0
            w = torch.randn(shp, generator=g) * (gain / math.sqrt(fan_in))
        out[k] = w.to(dtype)
    return out


# ----------------------------------------------------------------------------- forward

def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """get_timestep_embedding(flip_sin_to_cos=True, freq_shift=0) -> [cos | sin]
    (cf. gyre/pipeline/controlnet/models.py:166-173,467-472)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def resnet_block(P, p, x, temb, groups, eps):
    """ResnetBlock2D (block wiring cf. nonfree/tome_unet.py:82-112)."""
    h = F.silu(F.group_norm(x, groups, P[f"{p}.norm1.weight"], P[f"{p}.norm1.bias"], eps))
    h = F.conv2d(h, P[f"{p}.conv1.weight"], P[f"{p}.conv1.bias"], padding=1)
    if temb is not None:
        h = h + F.linear(F.silu(temb), P[f"{p}.time_emb_proj.weight"], P[f"{p}.time_emb_proj.bias"])[:, :, None, None]
    h = F.silu(F.group_norm(h, groups, P[f"{p}.norm2.weight"], P[f"{p}.norm2.bias"], eps))
    h = F.conv2d(h, P[f"{p}.conv2.weight"], P[f"{p}.conv2.bias"], padding=1)
    if f"{p}.conv_shortcut.weight" in P:
        x = F.conv2d(x, P[f"{p}.conv_shortcut.weight"], P[f"{p}.conv_shortcut.bias"])
    return x + h


ATTENTION_IMPL = "explicit"


def attention(P, p, x, ctx, heads, tome_r=0, upcast=False):
    """to_q/k/v -> [B*h,N,d] -> softmax(QK^T d^-1/2) V -> to_out
    (gyre/pipeline/models/memory_efficient_cross_attention.py:32-60).  With tome_r>0 this is
    ToMeMemoryEfficientCrossAttention.forward (nonfree/tome_memory_efficient_cross_attention.py:22-76):
    k and v are merged with ONE plan computed from k; q is untouched."""
    B, N, C = x.shape
    src = x if ctx is None else ctx
    q = F.linear(x, P[f"{p}.to_q.weight"])
    k = F.linear(src, P[f"{p}.to_k.weight"])
    v = F.linear(src, P[f"{p}.to_v.weight"])
    if tome_r > 0:
        plan = bipartite_soft_matching_plan(k, tome_r)
        if plan is not None:
            k = merge_mean(plan, k)
            v = merge_mean(plan, v)
    d = C // heads

    def split(t):
        return t.reshape(B, t.shape[1], heads, d).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    if ATTENTION_IMPL == "sdpa":      # bench.py --impl torchlib only: PyTorch's fused attention (same maths, library kernel)
        o = F.scaled_dot_product_attention(q, k, v)
    else:
        s = (q @ k.transpose(-1, -2)) * (d ** -0.5)
        o = torch.softmax(s, dim=-1) @ v
    o = o.permute(0, 2, 1, 3).reshape(B, N, C)
    return F.linear(o, P[f"{p}.to_out.0.weight"], P[f"{p}.to_out.0.bias"])


def transformer_2d(P, p, x, ctx, heads, groups, linear, tome_r=0):
    """Transformer2DModel + BasicTransformerBlock (cf. nonfree/tome_unet.py:114-136)."""
    B, C, H, W = x.shape
    res = x
    h = F.group_norm(x, groups, P[f"{p}.norm.weight"], P[f"{p}.norm.bias"], 1e-6)
    if linear:
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = F.linear(h, P[f"{p}.proj_in.weight"], P[f"{p}.proj_in.bias"])
    else:
        h = F.conv2d(h, P[f"{p}.proj_in.weight"], P[f"{p}.proj_in.bias"])
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    bi = 0
    while f"{p}.transformer_blocks.{bi}.norm1.weight" in P:
        b = f"{p}.transformer_blocks.{bi}"
        h = attention(P, f"{b}.attn1", F.layer_norm(h, (C,), P[f"{b}.norm1.weight"], P[f"{b}.norm1.bias"], 1e-5),
                      None, heads, tome_r) + h
        h = attention(P, f"{b}.attn2", F.layer_norm(h, (C,), P[f"{b}.norm2.weight"], P[f"{b}.norm2.bias"], 1e-5),
                      ctx, heads) + h
        n = F.layer_norm(h, (C,), P[f"{b}.norm3.weight"], P[f"{b}.norm3.bias"], 1e-5)
        a, g = F.linear(n, P[f"{b}.ff.net.0.proj.weight"], P[f"{b}.ff.net.0.proj.bias"]).chunk(2, dim=-1)
        h = F.linear(a * F.gelu(g), P[f"{b}.ff.net.2.weight"], P[f"{b}.ff.net.2.bias"]) + h
        bi += 1
    if linear:
        h = F.linear(h, P[f"{p}.proj_out.weight"], P[f"{p}.proj_out.bias"])
        h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    else:
        h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
        h = F.conv2d(h, P[f"{p}.proj_out.weight"], P[f"{p}.proj_out.bias"])
    return h + res


def num_transformer_blocks(cfg: UNetConfig) -> int:
    n = sum(cfg.layers_per_block for a in cfg.attn_levels if a) + 1
    n += sum(cfg.layers_per_block + 1 for a in cfg.attn_levels if a)
    return n


def unet_forward(P: dict, cfg: UNetConfig, sample, timestep, encoder_hidden_states, tome_r=0, taps=None,
                 added_cond_kwargs=None, down_block_additional_residuals=None, mid_block_additional_residual=None,
                 adapter_states=None):
    """UNet2DConditionModel.forward (call site gyre/pipeline/unet/core.py:274; encoder-half wiring
    cf. gyre/pipeline/controlnet/models.py:446-511; up path cf. nonfree/tome_unet.py:34-70).
    `tome_r`: int | (r, inflect) | list, expanded by parse_r over the transformer blocks in module
    execution order (nonfree/tome_unet.py:229-247).  `taps`: optional dict collecting named
    intermediate activations for per-layer parity tests."""
    B = sample.shape[0]
    ch = cfg.block_out_channels
    G, eps = cfg.norm_num_groups, cfg.norm_eps
    lin = cfg.use_linear_projection
    r_list = parse_r(num_transformer_blocks(cfg), tome_r) if tome_r else [0] * num_transformer_blocks(cfg)
    r_list = list(r_list)

    t = timestep
    if not torch.is_tensor(t):
        t = torch.tensor([t], dtype=torch.int64, device=sample.device)
    elif t.ndim == 0:
        t = t[None]
    t = t.expand(B)
    temb = timestep_embedding(t, ch[0]).to(sample.dtype)
    temb = F.linear(temb, P["time_embedding.linear_1.weight"], P["time_embedding.linear_1.bias"])
    temb = F.linear(F.silu(temb), P["time_embedding.linear_2.weight"], P["time_embedding.linear_2.bias"])
    if cfg.addition_time_embed_dim:
        # addition_embed_type == "text_time" (public diffusers UNet2DConditionModel): emb += add_embedding(cat(
        # [text_embeds, add_time_proj(time_ids.flatten()).reshape(B, -1)]))
        te, tid = added_cond_kwargs["text_embeds"], added_cond_kwargs["time_ids"]
        tproj = timestep_embedding(tid.flatten(), cfg.addition_time_embed_dim).reshape(B, -1)
        aug = torch.cat([te.to(sample.dtype), tproj.to(sample.dtype)], dim=-1)
        aug = F.linear(aug, P["add_embedding.linear_1.weight"], P["add_embedding.linear_1.bias"])
        aug = F.linear(F.silu(aug), P["add_embedding.linear_2.weight"], P["add_embedding.linear_2.bias"])
        temb = temb + aug

    def tap(name, v):
        if taps is not None:
            taps[name] = v

    h = F.conv2d(sample, P["conv_in.weight"], P["conv_in.bias"], padding=1)
    tap("conv_in", h)
    res = [h]
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block):
            h = resnet_block(P, f"down_blocks.{i}.resnets.{j}", h, temb, G, eps)
            tap(f"down_blocks.{i}.resnets.{j}", h)
            if cfg.attn_levels[i]:
                h = transformer_2d(P, f"down_blocks.{i}.attentions.{j}", h, encoder_hidden_states,
                                   cfg.num_heads[i], G, lin, r_list.pop(0))
                tap(f"down_blocks.{i}.attentions.{j}", h)
            res.append(h)
        if adapter_states:
            # T2I-adapter (gyre/pipeline/t2i_adapter/unet_patcher.py:21-60, 62-84): `hidden_states += adapter_state`
            # IN PLACE just before the block's downsampler (or after the block when it has none) - the tensor is
            # also the block's last entry of `output_states`, so that skip sees the state as well
            assert len(adapter_states) == len(ch)
            h = h + adapter_states[i].to(h.dtype)
            res[-1] = h
        if i < len(ch) - 1:
            h = F.conv2d(h, P[f"down_blocks.{i}.downsamplers.0.conv.weight"],
                         P[f"down_blocks.{i}.downsamplers.0.conv.bias"], stride=2, padding=1)
            res.append(h)
    h = resnet_block(P, "mid_block.resnets.0", h, temb, G, eps)
    h = transformer_2d(P, "mid_block.attentions.0", h, encoder_hidden_states, cfg.num_heads[-1], G, lin,
                       r_list.pop(0))
    h = resnet_block(P, "mid_block.resnets.1", h, temb, G, eps)
    # ControlNet residual injection as the caller expects it of UNet2DConditionModel.forward
    # (gyre/pipeline/unet/core.py:213-239 passes the summed ControlNet outputs through these two keywords):
    # the skips - not the down path - take the residuals, the mid block output takes its own
    if down_block_additional_residuals is not None:
        # (diffusers zips the two lists: a longer list of residuals - the 13 scalar zeros a cfg_only ControlNet hint returns
        # for the unconditional side, unified_pipeline.py:1001-1006 - is cut to the number of skips)
        assert len(down_block_additional_residuals) >= len(res)
        res = [r + a.to(r.dtype) for r, a in zip(res, down_block_additional_residuals)]
    if mid_block_additional_residual is not None:
        h = h + mid_block_additional_residual.to(h.dtype)
    tap("mid_block", h)
    rattn = list(reversed(cfg.attn_levels))
    rheads = list(reversed(cfg.num_heads))
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block + 1):
            h = torch.cat([h, res.pop()], dim=1)
            h = resnet_block(P, f"up_blocks.{i}.resnets.{j}", h, temb, G, eps)
            if rattn[i]:
                h = transformer_2d(P, f"up_blocks.{i}.attentions.{j}", h, encoder_hidden_states, rheads[i], G, lin,
                                   r_list.pop(0))
            tap(f"up_blocks.{i}.{j}", h)
        if i < len(ch) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, P[f"up_blocks.{i}.upsamplers.0.conv.weight"], P[f"up_blocks.{i}.upsamplers.0.conv.bias"],
                         padding=1)
    h = F.silu(F.group_norm(h, G, P["conv_norm_out.weight"], P["conv_norm_out.bias"], eps))
    return F.conv2d(h, P["conv_out.weight"], P["conv_out.bias"], padding=1)


class OracleUNet:
    """DiffusersUNet protocol object (gyre/pipeline/unet/types.py:30-39): returns `.sample`."""

    class _Out:
        def __init__(self, sample):
            self.sample = sample

    def __init__(self, cfg: UNetConfig, params: dict):
        self.config = cfg
        self.params = params
        self.r = 0  # ToMe: set like `unet.r = int(value)` (gyre/pipeline/unified_pipeline.py:1582-1584)

    def __call__(self, latents, t, *, encoder_hidden_states, added_cond_kwargs=None,
                 down_block_additional_residuals=None, mid_block_additional_residual=None, adapter_states=None, **_):
        with torch.no_grad():
            return self._Out(unet_forward(self.params, self.config, latents, t, encoder_hidden_states, self.r,
                                          added_cond_kwargs=added_cond_kwargs,
                                          down_block_additional_residuals=down_block_additional_residuals,
                                          mid_block_additional_residual=mid_block_additional_residual,
                                          adapter_states=adapter_states))
