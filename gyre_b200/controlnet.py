"""ControlNet encoder on the native kernels (reference: gyre/pipeline/controlnet/models.py:97-544,
called from gyre/pipeline/unet/core.py:96-239 which sums the outputs of all active ControlNets and passes them to the
UNet as `down_block_additional_residuals` / `mid_block_additional_residual`).

`B200ControlNet` has `ControlNetModel.forward`'s signature and output object; the residual tensors it returns are NCHW
fp16, exactly what `B200UNet(..., down_block_additional_residuals=, mid_block_additional_residual=)` takes."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _native as N
from .config import UNetConfig
from .weights import _depth, _resnet_keys, _transformer_keys

COND_CHANNELS = (16, 32, 96, 256)


def controlnet_param_shapes(cfg, conditioning_channels: int = 3) -> dict:
    """diffusers / gyre ControlNetModel state-dict names -> shapes."""
    cfg = UNetConfig.from_any(cfg)
    ch = cfg.block_out_channels
    T = ch[0] * 4
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
    skips, cin = [ch[0]], ch[0]
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


@dataclass
class ControlNetOutput:
    down_block_res_samples: tuple
    mid_block_res_sample: torch.Tensor


class B200ControlNet:
    def __init__(self, config, device=None, conditioning_channels: int = 3, channel_order: str = "rgb"):
        self.config = UNetConfig.from_any(config)
        if not torch.cuda.is_available():
            raise N.NativeError("B200ControlNet needs a CUDA device: there is no CPU path")
        if channel_order not in ("rgb", "bgr"):
            raise ValueError(f"unknown `controlnet_conditioning_channel_order`: {channel_order}")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = torch.float16
        self.channel_order = channel_order
        self.conditioning_channels = conditioning_channels
        self._lib = N.load()
        self._h = C.c_void_p()
        self._ws = {}
        self._loaded = False
        cfg = self.config
        c = N.UNetConfigC()
        c.in_channels, c.out_channels = cfg.in_channels, cfg.out_channels
        c.num_levels = len(cfg.block_out_channels)
        for i, v in enumerate(cfg.block_out_channels):
            c.block_out_channels[i] = v
            c.num_heads[i] = cfg.num_heads[i]
            c.attn_levels[i] = 1 if cfg.attn_levels[i] else 0
        c.layers_per_block = cfg.layers_per_block
        c.cross_attention_dim = cfg.cross_attention_dim
        c.norm_num_groups = cfg.norm_num_groups
        c.norm_eps = cfg.norm_eps
        c.use_linear_projection = int(cfg.use_linear_projection)
        c.upcast_attention = int(cfg.upcast_attention)
        for i, v in enumerate(cfg.transformer_layers_per_block[:len(cfg.block_out_channels)]):
            c.transformer_depth[i] = int(v)
        c.controlnet, c.conditioning_channels = 1, conditioning_channels
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_unet_create(C.byref(c), C.byref(self._h)), "unet_create")
        self.num_skips = self._lib.gyre_b200_unet_num_skips(self._h)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.gyre_b200_destroy(h)
            except Exception:
                pass
            self._h = None

    def load_state_dict(self, state_dict, strict: bool = True):
        keep = []
        with torch.cuda.device(self.device):
            try:
                for k, v in state_dict.items():
                    t = v.detach()
                    if t.dtype not in (torch.float16, torch.float32):
                        t = t.float()
                    t = t.to(self.device).contiguous()
                    keep.append(t)
                    shape = (C.c_int64 * t.ndim)(*t.shape)
                    N.check(self._lib.gyre_b200_load_weight(self._h, k.encode(), N.ptr(t), N.dtype_code(t), shape, t.ndim,
                                                            N.stream_ptr(self.device)), f"load_weight({k})")
            finally:
                torch.cuda.current_stream(self.device).synchronize()
                keep.clear()
        if strict:
            N.check(self._lib.gyre_b200_finalize(self._h), "finalize")
        self._loaded = True
        return self

    def _workspace(self, B, H, W, L):
        ws = self._ws.get((B, H, W, L))
        if ws is None:
            n = C.c_size_t()
            N.check(self._lib.gyre_b200_unet_workspace_bytes(self._h, B, H, W, L, C.byref(n)), "unet_workspace_bytes")
            self._ws.clear()
            ws = torch.empty((n.value,), device=self.device, dtype=torch.uint8)
            self._ws[(B, H, W, L)] = ws
        return ws

    def _skip_shapes(self, B, H, W):
        cfg = self.config
        shapes = [(B, cfg.block_out_channels[0], H, W)]
        h, w = H, W
        for i, c in enumerate(cfg.block_out_channels):
            shapes += [(B, c, h, w)] * cfg.layers_per_block
            if i < len(cfg.block_out_channels) - 1:
                h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
                shapes.append((B, c, h, w))
        return shapes, (B, cfg.block_out_channels[-1], h, w)

    @torch.no_grad()
    def __call__(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale: float = 1.0,
                 return_dict: bool = True, out=None, **kwargs):
        """`out`: optional preallocated fp16 tensors (the skip residuals, then the mid residual) to write into - e.g. the
        guided half of zero-filled full-batch tensors (gyre_b200.hints, cfg_only)."""
        if not self._loaded:
            raise N.NativeError("B200ControlNet: weights not loaded")
        unsupported = [k for k, v in kwargs.items() if v is not None]
        if unsupported:
            raise NotImplementedError(f"B200ControlNet: unsupported arguments {unsupported}")
        N.require_cuda(sample, encoder_hidden_states, controlnet_cond)
        B, _, H, W = sample.shape
        if tuple(controlnet_cond.shape[-2:]) != (8 * H, 8 * W):
            raise ValueError(f"controlnet_cond is {tuple(controlnet_cond.shape[-2:])}, expected {(8 * H, 8 * W)}")
        if self.channel_order == "bgr":
            controlnet_cond = torch.flip(controlnet_cond, dims=[1])
        x = sample.to(torch.float16).contiguous()
        cond = controlnet_cond.to(torch.float16).contiguous()
        ctx = encoder_hidden_states.to(torch.float16).contiguous()
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], device=self.device)
        if t.ndim == 0:
            t = t[None]
        t = t.to(device=self.device, dtype=torch.int64).expand(B).contiguous()
        L = ctx.shape[1]
        shapes, mid_shape = self._skip_shapes(B, H, W)
        if out is not None:
            if len(out) != len(shapes) + 1:
                raise ValueError(f"out: expected {len(shapes) + 1} tensors, got {len(out)}")
            for o, shp in zip(out, shapes + [mid_shape]):
                if tuple(o.shape) != tuple(shp) or o.dtype != torch.float16 or not o.is_contiguous():
                    raise ValueError(f"out: expected contiguous fp16 {tuple(shp)}, got {o.dtype} {tuple(o.shape)}")
            down, mid = list(out[:-1]), out[-1]
        else:
            down = [torch.empty(s, device=self.device, dtype=torch.float16) for s in shapes]
            mid = torch.empty(mid_shape, device=self.device, dtype=torch.float16)
        ptrs = (C.c_void_p * len(down))(*[d.data_ptr() for d in down])
        ws = self._workspace(B, H, W, L)
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_controlnet_forward(self._h, N.ptr(x), N.ptr(t), N.ptr(ctx), N.ptr(cond), B, H, W, L,
                                                           ptrs, len(down), N.ptr(mid), N.ptr(ws), ws.numel(),
                                                           N.stream_ptr(self.device)), "controlnet_forward")
        if conditioning_scale != 1.0:
            for d in down:
                d.mul_(conditioning_scale)
            mid.mul_(conditioning_scale)
        if not return_dict:
            return tuple(down), mid
        return ControlNetOutput(down_block_res_samples=tuple(down), mid_block_res_sample=mid)
