"""Oracle: diffusers-0.16 ``AutoencoderKL`` encode/decode restated functionally (TEST ONLY).

Third-party arithmetic (diffusers ~=0.16.0, not vendored).  Call sites in the reference:
gyre/pipeline/unified_pipeline.py:1523-1536 (decode), :305-318 (encode); hyper-parameters
gyre/ldm_config/v1-inference.yaml:46-67; op list SURVEY.md Appendix A / A.2.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: tuple = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215

    @staticmethod
    def sd():
        return VAEConfig()

    @staticmethod
    def tiny():
        return VAEConfig(block_out_channels=(64, 64, 128, 128))


def _res_keys(p, cin, cout):
    ks = {
        f"{p}.norm1.weight": (cin,), f"{p}.norm1.bias": (cin,),
        f"{p}.conv1.weight": (cout, cin, 3, 3), f"{p}.conv1.bias": (cout,),
        f"{p}.norm2.weight": (cout,), f"{p}.norm2.bias": (cout,),
        f"{p}.conv2.weight": (cout, cout, 3, 3), f"{p}.conv2.bias": (cout,),
    }
    if cin != cout:
        ks[f"{p}.conv_shortcut.weight"] = (cout, cin, 1, 1)
        ks[f"{p}.conv_shortcut.bias"] = (cout,)
    return ks


def _attn_keys(p, c):
    ks = {f"{p}.group_norm.weight": (c,), f"{p}.group_norm.bias": (c,)}
    for n in ("query", "key", "value", "proj_attn"):
        ks[f"{p}.{n}.weight"] = (c, c)
        ks[f"{p}.{n}.bias"] = (c,)
    return ks


def vae_param_shapes(cfg: VAEConfig, encoder=True, decoder=True) -> dict:
    ch = cfg.block_out_channels
    z = cfg.latent_channels
    ks = {}
    if decoder:
        top = ch[-1]
        ks["post_quant_conv.weight"] = (z, z, 1, 1)
        ks["post_quant_conv.bias"] = (z,)
        ks["decoder.conv_in.weight"] = (top, z, 3, 3)
        ks["decoder.conv_in.bias"] = (top,)
        ks.update(_res_keys("decoder.mid_block.resnets.0", top, top))
        ks.update(_attn_keys("decoder.mid_block.attentions.0", top))
        ks.update(_res_keys("decoder.mid_block.resnets.1", top, top))
        cin = top
        for i, c in enumerate(reversed(ch)):
            for j in range(cfg.layers_per_block + 1):
                ks.update(_res_keys(f"decoder.up_blocks.{i}.resnets.{j}", cin, c))
                cin = c
            if i < len(ch) - 1:
                ks[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (c, c, 3, 3)
                ks[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (c,)
        ks["decoder.conv_norm_out.weight"] = (ch[0],)
        ks["decoder.conv_norm_out.bias"] = (ch[0],)
        ks["decoder.conv_out.weight"] = (cfg.out_channels, ch[0], 3, 3)
        ks["decoder.conv_out.bias"] = (cfg.out_channels,)
    if encoder:
        ks["encoder.conv_in.weight"] = (ch[0], cfg.in_channels, 3, 3)
        ks["encoder.conv_in.bias"] = (ch[0],)
        cin = ch[0]
        for i, c in enumerate(ch):
            for j in range(cfg.layers_per_block):
                ks.update(_res_keys(f"encoder.down_blocks.{i}.resnets.{j}", cin, c))
                cin = c
            if i < len(ch) - 1:
                ks[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = (c, c, 3, 3)
                ks[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = (c,)
        ks.update(_res_keys("encoder.mid_block.resnets.0", cin, cin))
        ks.update(_attn_keys("encoder.mid_block.attentions.0", cin))
        ks.update(_res_keys("encoder.mid_block.resnets.1", cin, cin))
        ks["encoder.conv_norm_out.weight"] = (cin,)
        ks["encoder.conv_norm_out.bias"] = (cin,)
        ks["encoder.conv_out.weight"] = (2 * z, cin, 3, 3)
        ks["encoder.conv_out.bias"] = (2 * z,)
        ks["quant_conv.weight"] = (2 * z, 2 * z, 1, 1)
        ks["quant_conv.bias"] = (2 * z,)
    return ks


def _resnet(P, p, x, G):
    """ResnetBlock2D with temb=None, eps 1e-6."""
    h = F.silu(F.group_norm(x, G, P[f"{p}.norm1.weight"], P[f"{p}.norm1.bias"], 1e-6))
    h = F.conv2d(h, P[f"{p}.conv1.weight"], P[f"{p}.conv1.bias"], padding=1)
    h = F.silu(F.group_norm(h, G, P[f"{p}.norm2.weight"], P[f"{p}.norm2.bias"], 1e-6))
    h = F.conv2d(h, P[f"{p}.conv2.weight"], P[f"{p}.conv2.bias"], padding=1)
    if f"{p}.conv_shortcut.weight" in P:
        x = F.conv2d(x, P[f"{p}.conv_shortcut.weight"], P[f"{p}.conv_shortcut.bias"])
    return x + h


def _attn_block(P, p, x, G):
    """Legacy AttentionBlock: 1 head, d=C, Linear q/k/v with bias, fp32 softmax."""
    B, C, H, W = x.shape
    h = F.group_norm(x, G, P[f"{p}.group_norm.weight"], P[f"{p}.group_norm.bias"], 1e-6)
    h = h.view(B, C, H * W).transpose(1, 2)
    q = F.linear(h, P[f"{p}.query.weight"], P[f"{p}.query.bias"])
    k = F.linear(h, P[f"{p}.key.weight"], P[f"{p}.key.bias"])
    v = F.linear(h, P[f"{p}.value.weight"], P[f"{p}.value.bias"])
    s = (q @ k.transpose(-1, -2)) * (C ** -0.5)
    o = torch.softmax(s.float(), dim=-1).to(s.dtype) @ v
    o = F.linear(o, P[f"{p}.proj_attn.weight"], P[f"{p}.proj_attn.bias"])
    return o.transpose(-1, -2).reshape(B, C, H, W) + x


def vae_decode(P, cfg: VAEConfig, z, taps=None):
    """AutoencoderKL.decode(z).sample = Decoder(post_quant_conv(z))."""
    G = cfg.norm_num_groups
    ch = cfg.block_out_channels
    h = F.conv2d(z, P["post_quant_conv.weight"], P["post_quant_conv.bias"])
    h = F.conv2d(h, P["decoder.conv_in.weight"], P["decoder.conv_in.bias"], padding=1)
    h = _resnet(P, "decoder.mid_block.resnets.0", h, G)
    h = _attn_block(P, "decoder.mid_block.attentions.0", h, G)
    h = _resnet(P, "decoder.mid_block.resnets.1", h, G)
    if taps is not None:
        taps["mid"] = h
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block + 1):
            h = _resnet(P, f"decoder.up_blocks.{i}.resnets.{j}", h, G)
        if i < len(ch) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, P[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"],
                         P[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
        if taps is not None:
            taps[f"up{i}"] = h
    h = F.silu(F.group_norm(h, G, P["decoder.conv_norm_out.weight"], P["decoder.conv_norm_out.bias"], 1e-6))
    return F.conv2d(h, P["decoder.conv_out.weight"], P["decoder.conv_out.bias"], padding=1)


def vae_encode_moments(P, cfg: VAEConfig, x):
    """AutoencoderKL.encode(x) -> moments [B, 2z, H/8, W/8] (mean | logvar)."""
    G = cfg.norm_num_groups
    ch = cfg.block_out_channels
    h = F.conv2d(x, P["encoder.conv_in.weight"], P["encoder.conv_in.bias"], padding=1)
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block):
            h = _resnet(P, f"encoder.down_blocks.{i}.resnets.{j}", h, G)
        if i < len(ch) - 1:
            h = F.pad(h, (0, 1, 0, 1), value=0.0)   # Downsample2D with padding=0: asymmetric pad
            h = F.conv2d(h, P[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"],
                         P[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"], stride=2)
    h = _resnet(P, "encoder.mid_block.resnets.0", h, G)
    h = _attn_block(P, "encoder.mid_block.attentions.0", h, G)
    h = _resnet(P, "encoder.mid_block.resnets.1", h, G)
    h = F.silu(F.group_norm(h, G, P["encoder.conv_norm_out.weight"], P["encoder.conv_norm_out.bias"], 1e-6))
    h = F.conv2d(h, P["encoder.conv_out.weight"], P["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, P["quant_conv.weight"], P["quant_conv.bias"])


def gaussian_sample(moments, noise):
    """DiagonalGaussianDistribution.sample: mean + exp(0.5*clamp(logvar,-30,20)) * noise."""
    mean, logvar = moments.chunk(2, dim=1)
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise


class OracleVAE:
    """Duck-types what the pipeline touches: decode(x).sample, encode(x).latent_dist.sample(generator)."""

    class _Out:
        def __init__(self, sample):
            self.sample = sample

    def __init__(self, cfg: VAEConfig, params: dict, sample_dtype=None):
        self.config = cfg
        self.params = params
        self.sample_dtype = sample_dtype

    def decode(self, z):
        with torch.no_grad():
            return self._Out(vae_decode(self.params, self.config, z))

    class _Dist:
        """diffusers-0.16 DiagonalGaussianDistribution: logvar clamped to [-30, 20],
        sample = mean + std * randn(generator) drawn on the generator's device."""

        def __init__(self, moments, sample_dtype=None):
            self.mean, logvar = moments.chunk(2, dim=1)
            self.std = torch.exp(0.5 * logvar.clamp(-30.0, 20.0))
            # the draw happens in the VAE's dtype (fp16 on the GPU path, unified_pipeline.py:305-313): a
            # torch.randn in fp16 is a different stream from one in fp32, so the dtype is part of the seed contract
            self.sample_dtype = sample_dtype or self.mean.dtype

        def sample(self, generator=None):
            dev = generator.device if generator is not None else self.mean.device
            noise = torch.randn(self.mean.shape, generator=generator, device=dev, dtype=self.sample_dtype)
            return self.mean + self.std * noise.to(self.mean.device).to(self.mean.dtype)

    class _Enc:
        def __init__(self, dist):
            self.latent_dist = dist

    def encode(self, x):
        with torch.no_grad():
            return self._Enc(self._Dist(vae_encode_moments(self.params, self.config, x), self.sample_dtype))
