"""T2I-adapter encoder on the native kernels (reference: gyre/pipeline/t2i_adapter/adapter.py:65-132 `Adapter`,
configured by gyre/pipeline/t2i_adapter/models.py:80-123 `T2iAdapter_main`).  Its output is the `adapter_states` list the
UNet takes (gyre/pipeline/t2i_adapter/unet_patcher.py:21-60 adds state i to down block i's hidden state) - NCHW fp16
tensors, what `B200UNet(..., adapter_states=)` expects."""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N

MAIN_DEFAULTS = dict(cin=3 * 64, channels=(320, 640, 1280, 1280), nums_rb=2, ksize=1, sk=True, use_conv=False)


def adapter_param_shapes(channels=(320, 640, 1280, 1280), nums_rb=2, cin=192, ksize=1, sk=True, use_conv=False) -> dict:
    """nn.Module state-dict names of the reference's `Adapter` -> shapes."""
    ks = {"conv_in.weight": (channels[0], cin, 3, 3), "conv_in.bias": (channels[0],)}
    for i in range(len(channels)):
        for j in range(nums_rb):
            down = i != 0 and j == 0
            in_c, out_c = (channels[i - 1] if down else channels[i]), channels[i]
            p = f"body.{i * nums_rb + j}"
            if in_c != out_c or not sk:
                ks[f"{p}.in_conv.weight"] = (out_c, in_c, ksize, ksize)
                ks[f"{p}.in_conv.bias"] = (out_c,)
            ks[f"{p}.block1.weight"] = (out_c, out_c, 3, 3)
            ks[f"{p}.block1.bias"] = (out_c,)
            ks[f"{p}.block2.weight"] = (out_c, out_c, ksize, ksize)
            ks[f"{p}.block2.bias"] = (out_c,)
            if not sk:
                ks[f"{p}.skep.weight"] = (out_c, in_c, ksize, ksize)
                ks[f"{p}.skep.bias"] = (out_c,)
            if down and use_conv:
                ks[f"{p}.down_opt.op.weight"] = (in_c, in_c, 3, 3)
                ks[f"{p}.down_opt.op.bias"] = (in_c,)
    return ks


class B200T2iAdapter:
    """`T2iAdapter_main(channels, nums_rb, cin, ksize, sk, use_conv, autoinvert)`; called with the hint image
    [B, cin / 64, H, W] in [0, 1], returns the list of per-level feature maps."""

    def __init__(self, channels=(320, 640, 1280, 1280), nums_rb=3, cin=64, ksize=3, sk=False, use_conv=True,
                 autoinvert=False, device=None, light=False):
        if not torch.cuda.is_available():
            raise N.NativeError("B200T2iAdapter needs a CUDA device: there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = torch.float16
        self.channels, self.nums_rb, self.cin = tuple(channels), nums_rb, cin
        self.ksize, self.sk, self.use_conv, self.autoinvert = ksize, bool(sk), bool(use_conv), bool(autoinvert)
        self.is_light = bool(light)          # Adapter_light (adapter.py:240-263; `type: light`, models.py:186-195)
        self._lib = N.load()
        self._h = C.c_void_p()
        self._ws = {}
        self._loaded = False
        c = N.AdapterConfigC()
        c.cin, c.num_levels, c.nums_rb, c.ksize = cin, len(self.channels), nums_rb, ksize
        c.sk, c.use_conv, c.light = int(self.sk), int(self.use_conv), int(self.is_light)
        for i, v in enumerate(self.channels):
            c.channels[i] = v
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_adapter_create(C.byref(c), C.byref(self._h)), "adapter_create")

    @classmethod
    def light(cls, **config):
        """`type: light` adapters (T2iAdapter_light, models.py:186-195: cin 192, nums_rb 4): every level works at a quarter of
        the UNet's width - 1x1 in / out convolutions around plain conv-ReLU-conv residual blocks."""
        return cls(**{**dict(cin=192, channels=(320, 640, 1280, 1280), nums_rb=4), **config, "light": True})

    @classmethod
    def main(cls, **config):
        """The configuration gyre loads `type: main` adapters with (models.py:80-88 defaults merged with the engine's)."""
        return cls(**{**MAIN_DEFAULTS, **config})

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

    def _workspace(self, B, H, W):
        ws = self._ws.get((B, H, W))
        if ws is None:
            n = C.c_size_t()
            N.check(self._lib.gyre_b200_adapter_workspace_bytes(self._h, B, H, W, C.byref(n)), "adapter_workspace_bytes")
            self._ws.clear()
            ws = torch.empty((n.value,), device=self.device, dtype=torch.uint8)
            self._ws[(B, H, W)] = ws
        return ws

    @torch.no_grad()
    def __call__(self, x):
        if not self._loaded:
            raise N.NativeError("B200T2iAdapter: weights not loaded")
        N.require_cuda(x)
        B, Cimg, H, W = x.shape
        if Cimg * 64 != self.cin:
            raise ValueError(f"hint image has {Cimg} channels, the adapter takes {self.cin // 64}")
        if H % 8 or W % 8:
            raise ValueError(f"hint image {H}x{W} must be a multiple of 8")
        if self.autoinvert:
            # "If sample is more than 2/3 white, assume it needs inverting" (models.py:114-119) - without a host read
            x = torch.where(x.float().mean() > 0.66, 1 - x, x)
        x = x.to(torch.float16).contiguous()
        feats, h, w = [], H // 8, W // 8
        for i, c in enumerate(self.channels):
            if i:
                # (light adapters always pool: floor halving, like use_conv=False)
                h, w = ((h - 1) // 2 + 1, (w - 1) // 2 + 1) if (self.use_conv and not self.is_light) else (h // 2, w // 2)
            feats.append(torch.empty((B, c, h, w), device=self.device, dtype=torch.float16))
        ptrs = (C.c_void_p * len(feats))(*[f.data_ptr() for f in feats])
        ws = self._workspace(B, H, W)
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_adapter_forward(self._h, N.ptr(x), B, H, W, ptrs, len(feats), N.ptr(ws), ws.numel(),
                                                        N.stream_ptr(self.device)), "adapter_forward")
        return feats

    forward = __call__
