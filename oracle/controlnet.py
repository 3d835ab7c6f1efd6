"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's ControlNet encoder
(gyre/pipeline/controlnet/models.py: ControlNetConditioningEmbedding :41-94, ControlNetModel.__init__ :153-281,
ControlNetModel.forward :420-544).  The file itself imports diffusers (absent), but everything it wires is stated in it:
conv_in, `sample += controlnet_cond_embedding(cond)`, the UNet's down blocks and mid block (the blocks themselves are the
ones oracle/unet.py restates), then one zero-initialised 1x1 convolution per skip tensor and one for the mid output.
PINNED at what the file states: scripts/make_golden.py:pin_controlnet runs the reference's ControlNetModel (loaded by
scripts/_vendored.py:gyre_controlnet with the absent diffusers building blocks replaced by stand-ins that evaluate
oracle/unet.py's blocks) and asserts bit-equality of all 13 outputs, the conditioning embedding and the parameter inventory
(tests/golden/controlnet.pt).  PARITY UNPINNED below the block level (diffusers blocks), like oracle/unet.py.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .unet import (UNetConfig, _depth, _resnet_keys, _transformer_keys, resnet_block, timestep_embedding, transformer_2d)

COND_CHANNELS = (16, 32, 96, 256)       # conditioning_embedding_out_channels default (:129)


def controlnet_param_shapes(cfg: UNetConfig, conditioning_channels: int = 3) -> dict:
    ch = cfg.block_out_channels
    T = cfg.time_embed_dim
    ks = {
        "conv_in.weight": (ch[0], cfg.in_channels, 3, 3), "conv_in.bias": (ch[0],),
        "time_embedding.linear_1.weight": (T, ch[0]), "time_embedding.linear_1.bias": (T,),
        "time_embedding.linear_2.weight": (T, T), "time_embedding.linear_2.bias": (T,),
        "controlnet_cond_embedding.conv_in.weight": (COND_CHANNELS[0], conditioning_channels, 3, 3),
        "controlnet_cond_embedding.conv_in.bias": (COND_CHANNELS[0],),
        "controlnet_cond_embedding.conv_out.weight": (ch[0], COND_CHANNELS[-1], 3, 3),
        "controlnet_cond_embedding.conv_out.bias": (ch[0],),
    }
    for i in range(len(COND_CHANNELS) - 1):
        a, b = COND_CHANNELS[i], COND_CHANNELS[i + 1]
        ks[f"controlnet_cond_embedding.blocks.{2 * i}.weight"] = (a, a, 3, 3)
        ks[f"controlnet_cond_embedding.blocks.{2 * i}.bias"] = (a,)
        ks[f"controlnet_cond_embedding.blocks.{2 * i + 1}.weight"] = (b, a, 3, 3)
        ks[f"controlnet_cond_embedding.blocks.{2 * i + 1}.bias"] = (b,)
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
    for k, c in enumerate(skips):
        ks[f"controlnet_down_blocks.{k}.weight"] = (c, c, 1, 1)
        ks[f"controlnet_down_blocks.{k}.bias"] = (c,)
    ks["controlnet_mid_block.weight"] = (cin, cin, 1, 1)
    ks["controlnet_mid_block.bias"] = (cin,)
    return ks


def cond_embedding(P, cond):
    """ControlNetConditioningEmbedding.forward (:84-94): conv -> SiLU, six more convs each followed by SiLU (every second
    one stride 2), a zero-initialised conv_out without activation."""
    p = "controlnet_cond_embedding"
    e = F.silu(F.conv2d(cond, P[f"{p}.conv_in.weight"], P[f"{p}.conv_in.bias"], padding=1))
    for i in range(2 * (len(COND_CHANNELS) - 1)):
        e = F.silu(F.conv2d(e, P[f"{p}.blocks.{i}.weight"], P[f"{p}.blocks.{i}.bias"], padding=1, stride=1 + (i % 2)))
    return F.conv2d(e, P[f"{p}.conv_out.weight"], P[f"{p}.conv_out.bias"], padding=1)


def controlnet_forward(P, cfg: UNetConfig, sample, timestep, encoder_hidden_states, controlnet_cond):
    """ControlNetModel.forward (:420-544) -> (down_block_res_samples tuple, mid_block_res_sample)."""
    B = sample.shape[0]
    ch = cfg.block_out_channels
    G, eps, lin = cfg.norm_num_groups, cfg.norm_eps, cfg.use_linear_projection
    t = timestep
    if not torch.is_tensor(t):
        t = torch.tensor([t], dtype=torch.int64, device=sample.device)
    elif t.ndim == 0:
        t = t[None]
    t = t.expand(B)
    temb = timestep_embedding(t, ch[0]).to(sample.dtype)
    temb = F.linear(temb, P["time_embedding.linear_1.weight"], P["time_embedding.linear_1.bias"])
    temb = F.linear(F.silu(temb), P["time_embedding.linear_2.weight"], P["time_embedding.linear_2.bias"])
    h = F.conv2d(sample, P["conv_in.weight"], P["conv_in.bias"], padding=1)
    h = h + cond_embedding(P, controlnet_cond)
    res = [h]
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block):
            h = resnet_block(P, f"down_blocks.{i}.resnets.{j}", h, temb, G, eps)
            if cfg.attn_levels[i]:
                h = transformer_2d(P, f"down_blocks.{i}.attentions.{j}", h, encoder_hidden_states, cfg.num_heads[i], G, lin, 0)
            res.append(h)
        if i < len(ch) - 1:
            h = F.conv2d(h, P[f"down_blocks.{i}.downsamplers.0.conv.weight"], P[f"down_blocks.{i}.downsamplers.0.conv.bias"],
                         stride=2, padding=1)
            res.append(h)
    h = resnet_block(P, "mid_block.resnets.0", h, temb, G, eps)
    h = transformer_2d(P, "mid_block.attentions.0", h, encoder_hidden_states, cfg.num_heads[-1], G, lin, 0)
    h = resnet_block(P, "mid_block.resnets.1", h, temb, G, eps)
    down = tuple(F.conv2d(r, P[f"controlnet_down_blocks.{k}.weight"], P[f"controlnet_down_blocks.{k}.bias"])
                 for k, r in enumerate(res))
    mid = F.conv2d(h, P["controlnet_mid_block.weight"], P["controlnet_mid_block.bias"])
    return down, mid
