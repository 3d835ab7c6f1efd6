"""Host-side mirror of the UNet the reference pipeline drives: an object satisfying the `DiffusersUNet`
protocol (reference: gyre/pipeline/unet/types.py:30-39; sole call site gyre/pipeline/unet/core.py:274)
whose forward is ONE call into libgyre_b200 (gyre_b200_unet_forward).  PyTorch only owns the memory."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _native as N
from . import module_tree as MT
from .config import UNetConfig


@dataclass
class UNetOutput:
    """What `unet(...)` returns in the reference: an object with `.sample` (core.py:274)."""
    sample: torch.Tensor


def parse_r(num_layers: int, r):
    """ToMe r expansion (reference: nonfree/ToMe/tome/utils.py:80-105, called by ToMeUNet.forward,
    nonfree/tome_unet.py:229-247): int -> constant list; (r, inflect) -> linear ramp; list -> padded."""
    inflect = 0
    if isinstance(r, list):
        if len(r) < num_layers:
            r = r + [0] * (num_layers - len(r))
        return list(r)
    elif isinstance(r, tuple):
        r, inflect = r
    min_val = int(r * (1.0 - inflect))
    max_val = 2 * r - min_val
    step = (max_val - min_val) / (num_layers - 1)
    return [int(min_val + step * i) for i in range(num_layers)]


class B200UNet(torch.nn.Module):
    """UNet2DConditionModel replacement.  `unet(latents, t, encoder_hidden_states=ctx).sample`.

    An `nn.Module` with the surface the reference pipeline reads off its UNet: `.config.in_channels` /
    `.config.sample_size`, `.modules()` / `.named_modules()` / `.parameters()` under the diffusers names, `.dtype` /
    `.device`, `set_attention_slice`, `set_use_memory_efficient_attention_xformers`, ToMe's `.r`.  It HOLDS the original
    parameters next to the packed native copy (`from_module`: the caller's own module tree; `load_state_dict`: a tree
    rebuilt from the state dict), so gyre's LoRA hooks (gyre/pipeline/lora.py:99-166) attach to it like to the original
    - and are folded into the packed weights before the next forward (`_sync_lora`)."""

    def __init__(self, config, device=None, hold_parameters: bool = True):
        super().__init__()
        self.config = UNetConfig.from_any(config)
        if not torch.cuda.is_available():
            raise N.NativeError("B200UNet needs a CUDA device: there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = torch.float16
        self.r = 0                      # ToMe: `unet.r = int(value)` (unified_pipeline.py:1582-1584)
        self.hold_parameters = hold_parameters
        self._lib = N.load()
        self._h = C.c_void_p()
        self._ws = {}
        self._loaded = False
        self._ctx_bound = None          # (owner, B, L) of the context bound with set_context
        self._lora_sig = ()             # LoRA hooks folded into the packed weights right now
        self._attention_slice = None    # recorded only: the native attention is a flash kernel, nothing to slice
        self._xformers = True
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
        c.addition_embed_dim = int(cfg.projection_class_embeddings_input_dim if cfg.addition_time_embed_dim else 0)
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_unet_create(C.byref(c), C.byref(self._h)), "unet_create")
        self.num_transformer_blocks = self._lib.gyre_b200_unet_num_transformer_blocks(self._h)

    @classmethod
    def from_module(cls, module: torch.nn.Module, config=None, device=None):
        """Wraps the caller's original UNet (`diffusers.UNet2DConditionModel` layout): config from `module.config`, weights
        packed from `module.state_dict()`, and the module tree itself adopted - the parameters stay the caller's."""
        self = cls(config if config is not None else module.config, device=device, hold_parameters=False)
        MT.adopt_module(self, module)
        self._load_packed(module.state_dict(), strict=True)
        return self

    # -- the diffusers / gyre switches the pipeline flips on its UNet (unified_pipeline.py:1430-1450) -------------------
    def set_attention_slice(self, slice_size):
        self._attention_slice = slice_size

    def set_use_memory_efficient_attention_xformers(self, valid: bool, *args, **kwargs):
        self._xformers = bool(valid)

    def enable_xformers_memory_efficient_attention(self, *args, **kwargs):
        self._xformers = True

    def disable_xformers_memory_efficient_attention(self):
        self._xformers = False

    def to(self, *args, **kwargs):
        """The native weights are bound to `self.device`; `.to()` may only name that device (or a dtype for the held
        parameters, which does not affect the fp16 compute path)."""
        dev = kwargs.get("device", next((a for a in args if isinstance(a, (str, torch.device, int))), None))
        if dev is not None and torch.device(dev if not isinstance(dev, int) else f"cuda:{dev}").type == "cuda" and \
                torch.device(dev if not isinstance(dev, int) else f"cuda:{dev}").index not in (None, self.device.index):
            raise N.NativeError(f"B200UNet is bound to {self.device}; create a new one on {dev}")
        return self

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.gyre_b200_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- parameters ---------------------------------------------------------------------------------
    def load_weight(self, key: str, tensor: torch.Tensor, _keep=None):
        t = tensor.detach()
        if t.dtype not in (torch.float16, torch.float32):
            t = t.float()
        t = t.to(self.device).contiguous()
        shape = (C.c_int64 * t.ndim)(*t.shape)
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_load_weight(self._h, key.encode(), N.ptr(t), N.dtype_code(t), shape, t.ndim,
                                                    N.stream_ptr(self.device)), f"load_weight({key})")
            # the packing kernels read `t` asynchronously: the caller either keeps it alive until it synchronises
            # (load_state_dict: ONE synchronize for the whole state dict) or we wait here
            if _keep is not None:
                _keep.append(t)
            else:
                torch.cuda.current_stream(self.device).synchronize()

    def _load_packed(self, state_dict, strict=True):
        self.set_context(None)          # cached K/V projections depend on the attn2 weights
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
        self._lora_sig = ()

    def load_state_dict(self, state_dict, strict: bool = True):
        self._load_packed(state_dict, strict)
        if self.hold_parameters:
            MT.build_param_tree(self, state_dict)
        return self

    def _sync_lora(self):
        """Folds the LoRA hooks currently attached to the held modules into the packed weights (and un-folds the ones that
        were removed): `W + scale * alpha / r * up . down` per hooked layer (gyre/pipeline/lora.py:150-160 applies the
        same product to the layer's input at run time).  Only layers whose hooks changed are re-packed."""
        sig = MT.lora_signature(self)
        if sig == self._lora_sig:
            return
        touched = {e[0] for e in sig} | {e[0] for e in self._lora_sig}
        keep = []
        try:
            for key, w in MT.lora_folded_weights(self, sorted(touched)).items():
                self.load_weight(key, w, _keep=keep)
        finally:
            torch.cuda.current_stream(self.device).synchronize()
            keep.clear()
        self.set_context(None)
        self._lora_sig = sig

    # -- forward ------------------------------------------------------------------------------------
    def _workspace(self, B, H, W, L):
        key = (B, H, W, L)
        ws = self._ws.get(key)
        if ws is None:
            n = C.c_size_t()
            N.check(self._lib.gyre_b200_unet_workspace_bytes(self._h, B, H, W, L, C.byref(n)), "unet_workspace_bytes")
            # one live shape is the usual pattern; a hires-fix run alternates between two (natural / full size)
            while len(self._ws) >= 2:
                self._ws.pop(next(iter(self._ws)))
            ws = torch.empty((n.value,), device=self.device, dtype=torch.uint8)
            self._ws[key] = ws
        return ws

    def _timesteps(self, t, B):
        if not torch.is_tensor(t):
            t = torch.tensor([t], device=self.device)
        if t.ndim == 0:
            t = t[None]
        if t.is_floating_point():
            ti = t.round()
            if not torch.equal(ti, t):
                raise NotImplementedError("fractional timesteps are not supported (the reference quantises, "
                                          "common_scheduler.py:342-355)")
            t = ti
        return t.to(device=self.device, dtype=torch.int64).expand(B).contiguous()

    def tome_r_list(self):
        if not self.r:
            return None
        return parse_r(self.num_transformer_blocks, self.r)

    def set_context(self, ctx_f16, owner=None):
        """Binds `ctx` ([B, L, Cc] fp16) for the following `forward_raw(..., None)` calls: the cross-attention
        K/V projections are computed once here (gyre_b200_unet_set_context).  `None` drops the binding."""
        if ctx_f16 is None:
            if self._ctx_bound is not None:
                with torch.cuda.device(self.device):
                    N.check(self._lib.gyre_b200_unet_set_context(self._h, None, 0, 0, N.stream_ptr(self.device)),
                            "unet_set_context")
            self._ctx_bound = None
            return
        if not self._loaded:
            raise N.NativeError("B200UNet: weights not loaded")
        N.require_cuda(ctx_f16)
        B, L, _ = ctx_f16.shape
        with torch.cuda.device(self.device):
            N.check(self._lib.gyre_b200_unet_set_context(self._h, N.ptr(ctx_f16), B, L, N.stream_ptr(self.device)),
                    "unet_set_context")
        self._ctx_bound = (owner, B, L)

    def added_cond_vector(self, text_embeds, time_ids):
        """`addition_embed_type == "text_time"`: cat([text_embeds, sinusoid(time_ids).flatten(1)]) -> [B, proj_in] fp16.
        The sinusoid layout is the timestep embedding's ([cos | sin], flip_sin_to_cos, shift 0)."""
        import math
        dim = self.config.addition_time_embed_dim
        half = dim // 2
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=self.device) / half)
        ang = time_ids.to(self.device).float().reshape(-1)[:, None] * freqs[None]
        emb = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1).reshape(time_ids.shape[0], -1)
        return torch.cat([text_embeds.to(self.device).float(), emb], dim=-1).to(torch.float16).contiguous()

    def forward_raw(self, sample_f16, t_i64, ctx_f16, out=None, add_cond=None, cfg_duplicate: bool = False,
                    down_block_additional_residuals=None, mid_block_additional_residual=None, adapter_states=None):
        if down_block_additional_residuals is None and mid_block_additional_residual is None and not adapter_states:
            return self._forward_raw(sample_f16, t_i64, ctx_f16, out, add_cond, cfg_duplicate)
        keep = None
        try:
            keep = [self._bind_control_residuals(down_block_additional_residuals, mid_block_additional_residual, sample_f16),
                    self._bind_adapter_states(adapter_states, sample_f16)]
            return self._forward_raw(sample_f16, t_i64, ctx_f16, out, add_cond, cfg_duplicate)
        finally:
            # valid for ONE forward, like the keyword arguments (see forward)
            self._lib.gyre_b200_unet_set_control_residuals(self._h, None, 0, None)
            self._lib.gyre_b200_unet_set_adapter_states(self._h, None, 0)
            del keep

    def _forward_raw(self, sample_f16, t_i64, ctx_f16, out=None, add_cond=None, cfg_duplicate: bool = False):
        """No conversions: fp16 NCHW sample, int64 [B] timesteps, fp16 [B, L, Cc] context (or None: the context
        bound with `set_context`), all on device.  `cfg_duplicate`: the caller's promise that the batch is [x ; x] with
        equal timesteps in both halves (CFGUNet_Parallel): the part of the network in front of the first cross-attention
        then runs once for both halves (gyre_b200_unet_set_cfg_duplicate)."""
        if not self._loaded:
            raise N.NativeError("B200UNet: weights not loaded")
        if bool(cfg_duplicate) != getattr(self, "_cfg_dup", False):
            N.check(self._lib.gyre_b200_unet_set_cfg_duplicate(self._h, 1 if cfg_duplicate else 0), "unet_set_cfg_duplicate")
            self._cfg_dup = bool(cfg_duplicate)
        B, Cin, H, W = sample_f16.shape
        if ctx_f16 is None:
            if self._ctx_bound is None or self._ctx_bound[1] != B:
                raise N.NativeError("B200UNet: no context bound for this batch")
            L = self._ctx_bound[2]
        else:
            L = ctx_f16.shape[1]
        if out is None:
            out = torch.empty((B, self.config.out_channels, H, W), device=self.device, dtype=torch.float16)
        ws = self._workspace(B, H, W, L)
        r_list = self.tome_r_list()
        r_arr = (C.c_int32 * len(r_list))(*r_list) if r_list else None
        N.check(self._lib.gyre_b200_unet_forward_cond(self._h, N.ptr(sample_f16), N.ptr(t_i64), N.ptr(ctx_f16),
                                                      N.ptr(add_cond), B, H, W, L, r_arr, N.ptr(out), N.ptr(ws),
                                                      ws.numel(), N.stream_ptr(self.device)), "unet_forward")
        return out

    def _bind_control_residuals(self, down, mid, x):
        """ControlNet residuals (`down_block_additional_residuals` / `mid_block_additional_residual`, passed through by
        gyre/pipeline/unet/core.py:213-239) for the next native forward.  Returns the fp16 tensors to keep alive."""
        if down is None and mid is None:
            # nothing to bind: make sure no pointer of an earlier (possibly failed) forward is still set
            N.check(self._lib.gyre_b200_unet_set_control_residuals(self._h, None, 0, None), "unet_set_control_residuals")
            return None
        B, _, H, W = x.shape
        cfg = self.config
        keep, ptrs = [], None
        if down is not None:
            n = self._lib.gyre_b200_unet_num_skips(self._h)
            if len(down) != n:
                raise ValueError(f"expected {n} down_block_additional_residuals, got {len(down)}")
            chans, sizes = [cfg.block_out_channels[0]], [(H, W)]
            h, w = H, W
            for i, c in enumerate(cfg.block_out_channels):
                chans += [c] * cfg.layers_per_block
                sizes += [(h, w)] * cfg.layers_per_block
                if i < len(cfg.block_out_channels) - 1:
                    h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
                    chans.append(c)
                    sizes.append((h, w))
            for k, r in enumerate(down):
                if tuple(r.shape) != (B, chans[k], *sizes[k]):
                    raise ValueError(f"down residual {k}: expected {(B, chans[k], *sizes[k])}, got {tuple(r.shape)}")
                keep.append(r.to(device=x.device, dtype=torch.float16).contiguous())
            ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in keep])
        m = None
        if mid is not None:
            m = mid.to(device=x.device, dtype=torch.float16).contiguous()
            keep.append(m)
        N.check(self._lib.gyre_b200_unet_set_control_residuals(self._h, ptrs, len(down) if down is not None else 0,
                                                               N.ptr(m)), "unet_set_control_residuals")
        return keep

    def _bind_adapter_states(self, states, x):
        """T2I-adapter states (`adapter_states=`, gyre/pipeline/t2i_adapter/unet_patcher.py:95-110) for the next
        native forward: one tensor per down block.  Returns the fp16 tensors to keep alive."""
        if not states:                       # None or empty list: the reference's hook ignores both
            N.check(self._lib.gyre_b200_unet_set_adapter_states(self._h, None, 0), "unet_set_adapter_states")
            return None
        B, _, H, W = x.shape
        ch = self.config.block_out_channels
        if len(states) != len(ch):
            raise ValueError(f"expected {len(ch)} adapter_states (one per down block), got {len(states)}")
        keep = []
        h, w = H, W
        for i, s in enumerate(states):
            if tuple(s.shape) != (B, ch[i], h, w):
                raise ValueError(f"adapter state {i}: expected {(B, ch[i], h, w)}, got {tuple(s.shape)}")
            keep.append(s.to(device=x.device, dtype=torch.float16).contiguous())
            h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        ptrs = (C.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
        N.check(self._lib.gyre_b200_unet_set_adapter_states(self._h, ptrs, len(keep)), "unet_set_adapter_states")
        return keep

    def forward(self, latents, t, *, encoder_hidden_states, down_block_additional_residuals=None,
                 mid_block_additional_residual=None, adapter_states=None, added_cond_kwargs=None, **kwargs):
        N.require_cuda(latents, encoder_hidden_states)
        self._sync_lora()
        B = latents.shape[0]
        if latents.shape[1] != self.config.in_channels:
            raise ValueError(f"expected {self.config.in_channels} input channels, got {latents.shape[1]}")
        x = latents.to(torch.float16).contiguous()
        ctx = encoder_hidden_states.to(torch.float16).contiguous()
        if ctx.shape[0] != B:
            raise ValueError("encoder_hidden_states batch does not match latents")
        add = None
        if self.config.addition_time_embed_dim:
            if not added_cond_kwargs or "text_embeds" not in added_cond_kwargs or "time_ids" not in added_cond_kwargs:
                raise ValueError("this UNet needs added_cond_kwargs = {text_embeds, time_ids} (text_time conditioning)")
            add = self.added_cond_vector(added_cond_kwargs["text_embeds"], added_cond_kwargs["time_ids"])
        keep = None
        try:
            keep = [self._bind_control_residuals(down_block_additional_residuals, mid_block_additional_residual, x),
                    self._bind_adapter_states(adapter_states, x)]
            out = self.forward_raw(x, self._timesteps(t, B), ctx, add_cond=add)
        finally:
            # the residual / state pointers are valid for ONE forward, like the keyword arguments: whatever happened
            # above (a failed bind, a failed forward), nothing may stay bound once `keep` is released
            self._lib.gyre_b200_unet_set_control_residuals(self._h, None, 0, None)
            self._lib.gyre_b200_unet_set_adapter_states(self._h, None, 0)
            del keep
        return UNetOutput(sample=out.to(latents.dtype) if latents.dtype != torch.float16 else out)
