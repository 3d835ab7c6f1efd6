"""Parameter inventories (diffusers state-dict key -> shape) for the UNet / VAE this library runs, and a
deterministic synthetic initialiser.  No SD checkpoints exist on the build or bench machines (SURVEY.md
finding 1), so benchmarks and parity tests use seeded random weights of the exact architecture; a real
checkpoint loads through the same `load_state_dict` because the key names are diffusers' own
(SURVEY.md Appendix A)."""
from __future__ import annotations

import math

import torch

from .config import UNetConfig, VAEConfig


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
    proj = (c, c) if linear else (c, c, 1, 1)
    ks[f"{p}.proj_in.weight"] = proj
    ks[f"{p}.proj_out.weight"] = proj
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


def unet_param_shapes(cfg) -> dict:
    cfg = UNetConfig.from_any(cfg)
    ch = cfg.block_out_channels
    T = ch[0] * 4
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


def _vae_res_keys(p, cin, cout):
    return _resnet_keys(p, cin, cout, 0)


def _vae_attn_keys(p, c):
    ks = {f"{p}.group_norm.weight": (c,), f"{p}.group_norm.bias": (c,)}
    for n in ("query", "key", "value", "proj_attn"):
        ks[f"{p}.{n}.weight"] = (c, c)
        ks[f"{p}.{n}.bias"] = (c,)
    return ks


def vae_param_shapes(cfg) -> dict:
    cfg = VAEConfig.from_any(cfg)
    ch = cfg.block_out_channels
    z = cfg.latent_channels
    top = ch[-1]
    ks = {"post_quant_conv.weight": (z, z, 1, 1), "post_quant_conv.bias": (z,),
          "decoder.conv_in.weight": (top, z, 3, 3), "decoder.conv_in.bias": (top,)}
    ks.update(_vae_res_keys("decoder.mid_block.resnets.0", top, top))
    ks.update(_vae_attn_keys("decoder.mid_block.attentions.0", top))
    ks.update(_vae_res_keys("decoder.mid_block.resnets.1", top, top))
    cin = top
    for i, c in enumerate(reversed(ch)):
        for j in range(cfg.layers_per_block + 1):
            ks.update(_vae_res_keys(f"decoder.up_blocks.{i}.resnets.{j}", cin, c))
            cin = c
        if i < len(ch) - 1:
            ks[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (c, c, 3, 3)
            ks[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (c,)
    ks["decoder.conv_norm_out.weight"] = (ch[0],)
    ks["decoder.conv_norm_out.bias"] = (ch[0],)
    ks["decoder.conv_out.weight"] = (cfg.out_channels, ch[0], 3, 3)
    ks["decoder.conv_out.bias"] = (cfg.out_channels,)
    ks["encoder.conv_in.weight"] = (ch[0], cfg.in_channels, 3, 3)
    ks["encoder.conv_in.bias"] = (ch[0],)
    cin = ch[0]
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block):
            ks.update(_vae_res_keys(f"encoder.down_blocks.{i}.resnets.{j}", cin, c))
            cin = c
        if i < len(ch) - 1:
            ks[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = (c, c, 3, 3)
            ks[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = (c,)
    ks.update(_vae_res_keys("encoder.mid_block.resnets.0", cin, cin))
    ks.update(_vae_attn_keys("encoder.mid_block.attentions.0", cin))
    ks.update(_vae_res_keys("encoder.mid_block.resnets.1", cin, cin))
    ks["encoder.conv_norm_out.weight"] = (cin,)
    ks["encoder.conv_norm_out.bias"] = (cin,)
    ks["encoder.conv_out.weight"] = (2 * z, cin, 3, 3)
    ks["encoder.conv_out.bias"] = (2 * z,)
    ks["quant_conv.weight"] = (2 * z, 2 * z, 1, 1)
    ks["quant_conv.bias"] = (2 * z,)
    return ks


# residual-branch output layers get a reduced gain so that chained forwards of a random-init net stay O(1)
_BRANCH_OUT = ("conv2.weight", "to_out.0.weight", "ff.net.2.weight", "proj_out.weight", "proj_attn.weight")


def synth_state_dict(shapes: dict, seed: int, dtype=torch.float32, branch_gain: float = 0.5, device="cpu") -> dict:
    """W ~ N(0, gain^2 / fan_in), norm gamma ~ 1 + 0.1 N, biases ~ 0.02 N; ONE generator walked in sorted-key
    order.  device="cpu" reproduces the parity fixtures bit for bit; device="cuda" is the fast path for
    throughput runs where only the architecture matters."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for k in sorted(shapes):
        shp = shapes[k]
        if k.endswith(".weight") and len(shp) == 1:
            w = 1.0 + 0.1 * torch.randn(shp, generator=g, device=device)
        elif k.endswith(".bias"):
            w = 0.02 * torch.randn(shp, generator=g, device=device)
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            gain = branch_gain if k.endswith(_BRANCH_OUT) else 1.0
            w = torch.randn(shp, generator=g, device=device) * (gain / math.sqrt(fan_in))
        out[k] = w.to(dtype)
    return out
