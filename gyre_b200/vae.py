"""Host-side mirror of AutoencoderKL as the reference pipeline uses it: `vae.decode(x).sample`
(reference: gyre/pipeline/unified_pipeline.py:1523-1536) and
`vae.encode(img).latent_dist.sample(generator=g)` (:305-318)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _native as N
from . import module_tree as MT
from .config import VAEConfig


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class DiagonalGaussianDistribution:
    """mean | logvar split, logvar clamped to [-30, 20]; sample = mean + std * randn(generator)
    (diffusers-0.16 semantics; the draw happens on the generator's device)."""

    def __init__(self, moments: torch.Tensor):
        self.parameters = moments
        self.mean, logvar = moments.float().chunk(2, dim=1)
        self.logvar = logvar.clamp(-30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        dev = generator.device if generator is not None else self.mean.device
        noise = torch.randn(self.mean.shape, generator=generator, device=dev, dtype=self.parameters.dtype)
        return (self.mean + self.std * noise.to(self.mean.device).float()).to(self.parameters.dtype)

    def mode(self):
        return self.mean.to(self.parameters.dtype)


@dataclass
class EncoderOutput:
    latent_dist: DiagonalGaussianDistribution


class B200VAE(torch.nn.Module):
    """AutoencoderKL replacement: an `nn.Module` holding the original parameters under the diffusers names next to the
    packed native copy (see B200UNet), with the switches the reference flips on its VAE: `enable_slicing` /
    `disable_slicing` (decode one image per call - same results, the reference's low-memory mode,
    pipeline_wrapper.py:171-186), `enable_tiling` / `disable_tiling` (accepted and recorded; the native decoder never
    needs it - a 1024x1024 decode takes ~3.5 GB of workspace - and tiled decoding would CHANGE the image through its
    seam blending, so full-frame decoding is kept), `.dtype` (`vae_dtype`, unified_pipeline.py:1526-1536).

    `dtype=torch.float32` gives the fp32 INTERFACE the reference uses for VAEs that overflow in fp16 (inputs and outputs
    are fp32 tensors); the arithmetic stays fp16 tensor-core MMA with fp32 accumulation and fp32 GroupNorm statistics -
    a true fp32 convolution path is not built."""

    def __init__(self, config, device=None, dtype=torch.float16, hold_parameters: bool = True):
        super().__init__()
        self.config = VAEConfig.from_any(config)
        if not torch.cuda.is_available():
            raise N.NativeError("B200VAE needs a CUDA device: there is no CPU path")
        if dtype not in (torch.float16, torch.float32):
            raise ValueError("B200VAE dtype must be float16 or float32")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = dtype
        self.hold_parameters = hold_parameters
        self.use_slicing = False
        self.use_tiling = False
        self._lib = N.load()
        self._h = C.c_void_p()
        self._ws = {}
        self._loaded = False
        cfg = self.config
        c = N.VAEConfigC()
        c.in_channels, c.out_channels, c.latent_channels = cfg.in_channels, cfg.out_channels, cfg.latent_channels
        c.num_levels = len(cfg.block_out_channels)
        for i, v in enumerate(cfg.block_out_channels):
            c.block_out_channels[i] = v
        c.layers_per_block = cfg.layers_per_block
        c.norm_num_groups = cfg.norm_num_groups
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_vae_create(C.byref(c), C.byref(self._h)), "vae_create")

    @classmethod
    def from_module(cls, module: torch.nn.Module, config=None, device=None, dtype=None):
        p = next(module.parameters(), None)
        dt = dtype or (torch.float32 if p is not None and p.dtype == torch.float32 else torch.float16)
        self = cls(config if config is not None else module.config, device=device, dtype=dt, hold_parameters=False)
        MT.adopt_module(self, module)
        self._load_packed(module.state_dict(), strict=True)
        return self

    def enable_slicing(self):
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    def enable_tiling(self, *args, **kwargs):
        self.use_tiling = True

    def disable_tiling(self):
        self.use_tiling = False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.gyre_b200_destroy(h)
            except Exception:
                pass
            self._h = None

    def load_weight(self, key, tensor, _keep=None):
        t = tensor.detach()
        if t.dtype not in (torch.float16, torch.float32):
            t = t.float()
        t = t.to(self.device).contiguous()
        shape = (C.c_int64 * t.ndim)(*t.shape)
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_load_weight(self._h, key.encode(), N.ptr(t), N.dtype_code(t), shape, t.ndim,
                                                    N.stream_ptr(self.device)), f"load_weight({key})")
            # the packing kernels read `t` asynchronously (see B200UNet.load_weight)
            if _keep is not None:
                _keep.append(t)
            else:
                torch.cuda.current_stream(self.device).synchronize()

    def load_state_dict(self, state_dict, strict: bool = True):
        self._load_packed(state_dict, strict)
        if self.hold_parameters:
            MT.build_param_tree(self, state_dict)
        return self

    def _load_packed(self, state_dict, strict=True):
        keep = []
        try:
            for k, v in state_dict.items():
                self.load_weight(k, v, _keep=keep)
        finally:
            torch.cuda.current_stream(self.device).synchronize()
            keep.clear()
        if strict:
            N.check(self._lib.gyre_b200_finalize(self._h), "finalize")
        self._loaded = True

    def _workspace(self, B, h, w):
        key = (B, h, w)
        ws = self._ws.get(key)
        if ws is None:
            n = C.c_size_t()
            N.check(self._lib.gyre_b200_vae_workspace_bytes(self._h, B, h, w, C.byref(n)), "vae_workspace_bytes")
            self._ws.clear()
            ws = torch.empty((n.value,), device=self.device, dtype=torch.uint8)
            self._ws[key] = ws
        return ws

    def decode_raw(self, z_f16, postprocess=False, want_u8=False):
        if not self._loaded:
            raise N.NativeError("B200VAE: weights not loaded")
        B, _, h, w = z_f16.shape
        img = torch.empty((B, 3, 8 * h, 8 * w), device=self.device, dtype=torch.float16)
        u8 = torch.empty((B, 8 * h, 8 * w, 3), device=self.device, dtype=torch.uint8) if want_u8 else None
        ws = self._workspace(B, h, w)
        N.check(self._lib.gyre_b200_vae_decode(self._h, N.ptr(z_f16), B, h, w, 1 if postprocess else 0, N.ptr(img),
                                               N.ptr(u8), N.ptr(ws), ws.numel(), N.stream_ptr(self.device)),
                "vae_decode")
        return img, u8

    def decode(self, z):
        N.require_cuda(z)
        z16 = z.to(torch.float16).contiguous()
        if self.use_slicing and z16.shape[0] > 1:
            img = torch.cat([self.decode_raw(z16[i:i + 1])[0] for i in range(z16.shape[0])])
        else:
            img, _ = self.decode_raw(z16)
        out_dtype = z.dtype if z.dtype in (torch.float16, torch.float32) else self.dtype
        return DecoderOutput(sample=img if out_dtype == torch.float16 else img.to(out_dtype))

    def encode(self, image):
        if not self._loaded:
            raise N.NativeError("B200VAE: weights not loaded")
        N.require_cuda(image)
        B, _, H, W = image.shape
        if H % 8 or W % 8:
            raise ValueError("image size must be a multiple of 8")
        x = image.to(torch.float16).contiguous()
        mom = torch.empty((B, 2 * self.config.latent_channels, H // 8, W // 8), device=self.device,
                          dtype=torch.float16)
        ws = self._workspace(B, H // 8, W // 8)
        N.check(self._lib.gyre_b200_vae_encode(self._h, N.ptr(x), B, H, W, N.ptr(mom), N.ptr(ws), ws.numel(),
                                               N.stream_ptr(self.device)), "vae_encode")
        return EncoderOutput(latent_dist=DiagonalGaussianDistribution(mom.to(image.dtype)))
