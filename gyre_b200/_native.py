"""ctypes binding of libgyre_b200.so (include/gyre_b200.h).  PyTorch is plumbing here: it owns device
memory and the stream; every arithmetic step goes through the C ABI.  There is NO fallback: a missing
library, a failing build or a non-zero status raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libgyre_b200.so")
_lock = threading.Lock()
_lib = None


class Epilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p), ("rowgroup_bias", C.c_void_p), ("rows_per_group", C.c_int32), ("rgb_ld", C.c_int32),
        ("residual", C.c_void_p), ("ldr", C.c_int32), ("act", C.c_int32), ("out_f32", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int32),
        ("sk_ws", C.c_void_p), ("sk_ws_bytes", C.c_size_t), ("sk_flags", C.c_void_p), ("sk_flags_count", C.c_int32),
        ("rowstat_out", C.c_void_p), ("ln_rowstat", C.c_void_p), ("ln_colsum", C.c_void_p),
        ("ln_parts", C.c_int32), ("ln_inv_c", C.c_float), ("ln_eps", C.c_float),
        ("gn_out", C.c_void_p), ("gn_groups", C.c_int32), ("gn_nparts", C.c_int32),
    ]


class Step(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("v_pred", C.c_int32), ("cfg", C.c_int32), ("guidance", C.c_float),
        ("sigma", C.c_float), ("c_in_next", C.c_float), ("dt", C.c_float), ("sigma_up", C.c_float),
        ("sqrt_a_t", C.c_float), ("sqrt_1m_a_t", C.c_float), ("sqrt_a_prev", C.c_float), ("dir_coef", C.c_float),
        ("noise_coef", C.c_float),
    ]


class UNetConfigC(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int32), ("out_channels", C.c_int32), ("num_levels", C.c_int32),
        ("block_out_channels", C.c_int32 * 4), ("num_heads", C.c_int32 * 4), ("attn_levels", C.c_int32 * 4),
        ("layers_per_block", C.c_int32), ("cross_attention_dim", C.c_int32), ("norm_num_groups", C.c_int32),
        ("norm_eps", C.c_float), ("use_linear_projection", C.c_int32), ("upcast_attention", C.c_int32),
        ("transformer_depth", C.c_int32 * 4), ("addition_embed_dim", C.c_int32),
        ("controlnet", C.c_int32), ("conditioning_channels", C.c_int32),
    ]


class ClipVisionConfigC(C.Structure):
    _fields_ = [
        ("image_size", C.c_int32), ("patch_size", C.c_int32), ("hidden_size", C.c_int32), ("intermediate_size", C.c_int32),
        ("num_layers", C.c_int32), ("num_heads", C.c_int32), ("hidden_act", C.c_int32), ("layer_norm_eps", C.c_float),
        ("projection_dim", C.c_int32), ("num_concepts", C.c_int32), ("num_special", C.c_int32),
    ]


class StyleAdapterConfigC(C.Structure):
    _fields_ = [("width", C.c_int32), ("context_dim", C.c_int32), ("num_head", C.c_int32), ("n_layers", C.c_int32),
                ("num_token", C.c_int32)]


class AdapterConfigC(C.Structure):
    _fields_ = [
        ("cin", C.c_int32), ("num_levels", C.c_int32), ("channels", C.c_int32 * 4), ("nums_rb", C.c_int32),
        ("ksize", C.c_int32), ("sk", C.c_int32), ("use_conv", C.c_int32), ("light", C.c_int32),
    ]


class ClipConfigC(C.Structure):
    _fields_ = [
        ("vocab_size", C.c_int32), ("hidden_size", C.c_int32), ("intermediate_size", C.c_int32),
        ("num_layers", C.c_int32), ("num_heads", C.c_int32), ("max_positions", C.c_int32), ("hidden_act", C.c_int32),
        ("layer_norm_eps", C.c_float),
    ]


class VAEConfigC(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int32), ("out_channels", C.c_int32), ("latent_channels", C.c_int32),
        ("num_levels", C.c_int32), ("block_out_channels", C.c_int32 * 4), ("layers_per_block", C.c_int32),
        ("norm_num_groups", C.c_int32),
    ]


_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/gyre_b200.h one to one
SIGNATURES = {
    "gyre_b200_abi_version": (_i, []),
    "gyre_b200_last_error": (_i, [C.c_char_p, _sz]),
    "gyre_b200_launch_count": (C.c_ulonglong, []),
    "gyre_b200_prof_enable": (_i, [_i]),
    "gyre_b200_prof_reset": (_i, []),
    "gyre_b200_debug_attention_trace": (_i, [_vp, _i]),
    "gyre_b200_prof_read": (_i, [_i, C.POINTER(C.c_ulonglong), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                C.POINTER(C.c_double)]),
    "gyre_b200_prof_read_roofline": (_i, [_i, C.c_double, C.c_double, C.POINTER(C.c_double)]),
    "gyre_b200_debug_mma_bench": (_i, [_i, _i, _i, _i, _i, _vp, _vp]),
    "gyre_b200_set_tunable": (_i, [C.c_char_p, _i]),
    "gyre_b200_get_tunable": (_i, [C.c_char_p, C.POINTER(_i)]),
    "gyre_b200_unet_create": (_i, [C.POINTER(UNetConfigC), C.POINTER(_vp)]),
    "gyre_b200_unet_num_transformer_blocks": (_i, [_vp]),
    "gyre_b200_load_weight": (_i, [_vp, C.c_char_p, _vp, _i, C.POINTER(_i64), _i, _vp]),
    "gyre_b200_finalize": (_i, [_vp]),
    "gyre_b200_unet_workspace_bytes": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_sz)]),
    "gyre_b200_unet_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, C.POINTER(C.c_int32), _vp, _vp, _sz, _vp]),
    "gyre_b200_unet_set_context": (_i, [_vp, _vp, _i, _i, _vp]),
    "gyre_b200_unet_set_cfg_duplicate": (_i, [_vp, _i]),
    "gyre_b200_unet_set_control_residuals": (_i, [_vp, C.POINTER(C.c_void_p), _i, _vp]),
    "gyre_b200_unet_num_skips": (_i, [_vp]),
    "gyre_b200_unet_set_adapter_states": (_i, [_vp, C.POINTER(C.c_void_p), _i]),
    "gyre_b200_unet_forward_cond": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, C.POINTER(C.c_int32), _vp, _vp, _sz, _vp]),
    "gyre_b200_controlnet_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, C.POINTER(C.c_void_p), _i, _vp, _vp, _sz, _vp]),
    "gyre_b200_resample_f32": (_i, [_vp, _i64, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "gyre_b200_webp_sizes": (_i, [_i, _i, _i, _i, C.POINTER(_sz), C.POINTER(_sz)]),
    "gyre_b200_webp_encode": (_i, [_vp, _i, _i, _i, _i, _vp, _sz, _vp, _vp, _sz, _vp]),
    "gyre_b200_png_sizes": (_i, [_i, _i, _i, _i, C.POINTER(_sz), C.POINTER(_sz)]),
    "gyre_b200_png_encode": (_i, [_vp, _i, _i, _i, _i, _vp, _sz, _vp, _vp, _sz, _vp]),
    "gyre_b200_clip_vision_create": (_i, [C.POINTER(ClipVisionConfigC), C.POINTER(_vp)]),
    "gyre_b200_clip_vision_workspace_bytes": (_i, [_vp, _i, C.POINTER(_sz)]),
    "gyre_b200_safety_scores": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "gyre_b200_clip_vision_hidden": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "gyre_b200_style_adapter_create": (_i, [C.POINTER(StyleAdapterConfigC), C.POINTER(_vp)]),
    "gyre_b200_style_adapter_workspace_bytes": (_i, [_vp, _i, _i, C.POINTER(_sz)]),
    "gyre_b200_style_adapter_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "gyre_b200_resample_u8": (_i, [_vp, _i64, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "gyre_b200_clip_normalize": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp, _vp]),
    "gyre_b200_adapter_create": (_i, [C.POINTER(AdapterConfigC), C.POINTER(_vp)]),
    "gyre_b200_adapter_workspace_bytes": (_i, [_vp, _i, _i, _i, C.POINTER(_sz)]),
    "gyre_b200_adapter_forward": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(C.c_void_p), _i, _vp, _sz, _vp]),
    "gyre_b200_vae_create": (_i, [C.POINTER(VAEConfigC), C.POINTER(_vp)]),
    "gyre_b200_vae_workspace_bytes": (_i, [_vp, _i, _i, _i, C.POINTER(_sz)]),
    "gyre_b200_vae_decode": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "gyre_b200_vae_encode": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "gyre_b200_clip_create": (_i, [C.POINTER(ClipConfigC), C.POINTER(_vp)]),
    "gyre_b200_clip_workspace_bytes": (_i, [_vp, _i, _i, C.POINTER(_sz)]),
    "gyre_b200_clip_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "gyre_b200_destroy": (_i, [_vp]),
    "gyre_b200_sched_step": (_i, [C.POINTER(Step), _vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _vp]),
    "gyre_b200_cfg_combine": (_i, [_vp, _f, _i, _i64, _vp, _vp, _vp]),
    "gyre_b200_denoise": (_i, [_vp, _vp, _i, _f, _f, _f, _i, _i64, _vp, _vp]),
    "gyre_b200_lincomb": (_i, [_i, C.POINTER(_vp), C.POINTER(_f), _i, _i64, _vp, _vp, _f, _i, _vp]),
    "gyre_b200_dpm_error_partials": (_i, [_vp, _vp, _vp, _f, _f, _i64, _vp, _vp]),
    "gyre_b200_dpm_error_num_partials": (_i, []),
    "gyre_b200_denoise_blend": (_i, [_vp, _vp, _i, _f, _f, _f, _i, _i64, _vp, _vp, _vp, _f, _vp]),
    "gyre_b200_sched_step_blend": (_i, [C.POINTER(Step), _vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _vp, _vp, _f, _vp]),
    "gyre_b200_cat_channels": (_i, [_vp, _i, _vp, _i, _i, _i, _i64, _vp, _vp]),
    "gyre_b200_scale_latents": (_i, [_vp, _f, _i, _i, _i64, _vp, _vp]),
    "gyre_b200_outpaint_scratch_bytes": (_sz, []),
    "gyre_b200_outpaint_match_histograms": (_i, [_vp, _vp, _vp, _i, _i64, _vp, _vp, _vp]),
    "gyre_b200_lpw_weight": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "gyre_b200_gemm_rowstat_parts": (_i, [_i, _i]),
    "gyre_b200_ln_finalize_rows": (_i, [_vp, _i, _i, _i, _f, _vp, _vp]),
    "gyre_b200_ln_fold_linear": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gyre_b200_resample_select": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _i,
                                       _vp, _i, _i, _i, _i, _vp]),
    "gyre_b200_rand_select": (_i, [_vp, _vp, _vp, _f, _i64, _vp, _vp]),
    "gyre_b200_tome_workspace_bytes": (_i, [_i, _i, _i, C.POINTER(_sz)]),
    "gyre_b200_tome_plan_offsets": (_i, [_i, _i, _i, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "gyre_b200_tome_merge_kv": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "gyre_b200_gemm": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _i, C.POINTER(Epilogue), _vp]),
    "gyre_b200_pack_geglu": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "gyre_b200_conv3x3": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, C.POINTER(Epilogue), _vp]),
    "gyre_b200_conv3x3_packed_elems": (_sz, [_i, _i]),
    "gyre_b200_pack_conv3x3": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "gyre_b200_upconv3x3_packed_elems": (_sz, [_i, _i]),
    "gyre_b200_pack_upconv3x3": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "gyre_b200_upconv2x": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, C.POINTER(Epilogue), _vp]),
    "gyre_b200_groupnorm_scratch_floats": (_sz, [_i, _i, _i]),
    "gyre_b200_groupnorm": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _f, _vp, _vp, _i, _vp, _vp, _vp]),
    "gyre_b200_conv3x3_gn_parts": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "gyre_b200_groupnorm_pre_ok": (_i, [_i, _i, _i]),
    "gyre_b200_groupnorm_pre": (_i, [_vp, _i, _i, _i, _i, _f, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    "gyre_b200_layernorm": (_i, [_vp, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "gyre_b200_attention": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _i, _vp]),
}


class NativeError(RuntimeError):
    """Non-zero status from libgyre_b200 (the reference surfaces failures as Python exceptions that
    gyre/services/exception_to_grpc.py maps to gRPC codes; we keep that contract)."""


def lib_path() -> str:
    return _LIB_PATH


def load(build_if_missing: bool = True):
    """Loads (building first if the .so is absent) and type-annotates the library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            if not build_if_missing:
                raise NativeError(f"{_LIB_PATH} is missing; run `python -m gyre_b200.build`")
            from . import build as _build
            _build.build()
        lib = C.CDLL(_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)     # AttributeError here == header/library drift: fail loudly
            fn.restype = res
            fn.argtypes = args
        if lib.gyre_b200_abi_version() != 2:
            raise NativeError("libgyre_b200 ABI version mismatch")
        _lib = lib
        return lib


def last_error() -> str:
    buf = C.create_string_buffer(2048)
    load().gyre_b200_last_error(buf, 2048)
    return buf.value.decode(errors="replace")


def check(status: int, what: str):
    if status != 0:
        raise NativeError(f"{what} failed ({status}): {last_error()}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    if t is None:
        return None
    return t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NativeError("gyre_b200 kernels need CUDA tensors (there is no CPU path)")


_DT = {torch.float16: 0, torch.float32: 1}


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise NativeError(f"unsupported parameter dtype {t.dtype}") from None


# ----------------------------------------------------------------------------------------------------
# building-block wrappers (used by the per-kernel parity tests and the attention patcher)

_sk_scratch = {}


def stream_k_scratch(device):
    """Stream-K scratch (partial accumulators + flags) for the standalone building-block calls, one per
    (device, stream): two streams running GEMMs concurrently must not share partial slots or flags (a model
    owns its own scratch; a handle is driven by one stream at a time, like the reference's pipeline slots)."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    if key not in _sk_scratch:
        _sk_scratch[key] = (torch.empty((48 << 20,), device=device, dtype=torch.uint8),
                            torch.zeros((256,), device=device, dtype=torch.int32))
    return _sk_scratch[key]


def _epilogue(out, bias=None, residual=None, act=0, rowgroup_bias=None, rows_per_group=1):
    e = Epilogue()
    ws, flags = stream_k_scratch(out.device)
    e.sk_ws, e.sk_ws_bytes, e.sk_flags, e.sk_flags_count = ptr(ws), ws.numel(), ptr(flags), flags.numel()
    e.bias = ptr(bias)
    e.rowgroup_bias = ptr(rowgroup_bias)
    e.rows_per_group = rows_per_group
    e.rgb_ld = rowgroup_bias.stride(0) if rowgroup_bias is not None else 0
    e.residual = ptr(residual)
    e.ldr = residual.stride(0) if residual is not None else 0
    e.act = act
    e.out_f32 = 1 if out.dtype == torch.float32 else 0
    e.out = ptr(out)
    e.ldo = out.stride(0)
    return e


def gemm(a, w, bias=None, residual=None, act=0, a2=None, out_dtype=torch.float16, rowgroup_bias=None,
         rows_per_group=1, out=None, rowstat_out=None, ln_rowstat=None, ln_colsum=None, ln_raw_parts=False, ln_eps=1e-5):
    """out = epilogue([a | a2] @ w.T); a [M, K1] fp16, w [N, K1+K2] fp16 (GEGLU: pre-packed), bias fp32.
    rowstat_out / ln_rowstat / ln_colsum: the folded-LayerNorm epilogues (include/gyre_b200.h)."""
    require_cuda(a, w)
    M, K1 = a.shape
    K2 = a2.shape[1] if a2 is not None else 0
    N = w.shape[0]
    n_out = N // 2 if act == 1 else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=out_dtype)
    e = _epilogue(out, bias, residual, act, rowgroup_bias, rows_per_group)
    e.rowstat_out, e.ln_rowstat, e.ln_colsum = ptr(rowstat_out), ptr(ln_rowstat), ptr(ln_colsum)
    if ln_raw_parts:       # ln_rowstat = the producer's partials [parts <= 8, M, 2]: folded by the consumer itself
        e.ln_parts, e.ln_inv_c, e.ln_eps = ln_rowstat.shape[0], 1.0 / K1, float(ln_eps)
    check(load().gyre_b200_gemm(ptr(a), a.stride(0), K1, ptr(a2), a2.stride(0) if a2 is not None else 0, K2, ptr(w),
                                w.stride(0), M, N, C.byref(e), stream_ptr(a.device)), "gemm")
    return out


def gemm_rowstat_buffer(M, N, device):
    """float2 [parts][M] scratch for `gemm(..., rowstat_out=)` with N output columns."""
    parts = load().gyre_b200_gemm_rowstat_parts(M, N)
    return torch.zeros((parts, M, 2), device=device, dtype=torch.float32)


def ln_finalize_rows(parts, C_cols, eps=1e-5):
    """(mean, rstd) [M, 2] fp32 from the row-segment partials a GEMM left."""
    nparts, M, _ = parts.shape
    out = torch.empty((M, 2), device=parts.device, dtype=torch.float32)
    check(load().gyre_b200_ln_finalize_rows(ptr(parts), nparts, M, C_cols, float(eps), ptr(out), stream_ptr(parts.device)),
          "ln_finalize_rows")
    return out


def ln_fold_linear(w, gamma, beta, bias=None):
    """(W' fp16 [N, K], colsum fp32 [N], lnbias fp32 [N]) for LayerNorm(gamma, beta) followed by Linear(w, bias)."""
    require_cuda(w, gamma, beta)
    w = w.contiguous()
    Nn, K = w.shape
    wo = torch.empty_like(w)
    cs = torch.empty((Nn,), device=w.device, dtype=torch.float32)
    lb = torch.empty((Nn,), device=w.device, dtype=torch.float32)
    check(load().gyre_b200_ln_fold_linear(ptr(w), Nn, K, ptr(gamma.float().contiguous()), ptr(beta.float().contiguous()),
                                          ptr(bias), ptr(wo), ptr(cs), ptr(lb), stream_ptr(w.device)), "ln_fold_linear")
    return wo, cs, lb


def pack_geglu(w, bias):
    require_cuda(w)
    F2, K = w.shape
    wp = torch.empty((F2, K), device=w.device, dtype=torch.float16)
    bp = torch.empty((F2,), device=w.device, dtype=torch.float32) if bias is not None else None
    w = w.contiguous()
    if bias is not None:
        bias = bias.contiguous()
    check(load().gyre_b200_pack_geglu(ptr(w), dtype_code(w), F2 // 2, K, ptr(bias),
                                      dtype_code(bias) if bias is not None else 0, ptr(wp), ptr(bp),
                                      stream_ptr(w.device)), "pack_geglu")
    return wp, bp


def pack_conv3x3(w):
    require_cuda(w)
    cout, cin = w.shape[0], w.shape[1]
    n = load().gyre_b200_conv3x3_packed_elems(cin, cout)
    wp = torch.empty((n,), device=w.device, dtype=torch.float16)
    w = w.contiguous()
    check(load().gyre_b200_pack_conv3x3(ptr(w), dtype_code(w), cin, cout, ptr(wp), stream_ptr(w.device)), "pack_conv3x3")
    return wp


def conv3x3(x_nhwc, wp, cout, bias=None, residual=None, stride=1, pad=1, act=0, rowgroup_bias=None,
            rows_per_group=1, gn_groups=0, gn_guard=False):
    """x [B, H, W, Cin] fp16 NHWC -> [B, Ho, Wo, Cout] fp16 NHWC.  gn_groups > 0: also returns the GroupNorm partials
    [B, nparts, gn_groups, 2] fp32 the epilogue left for `groupnorm_pre` (None when the shape cannot produce them)."""
    require_cuda(x_nhwc, wp)
    B, H, W, Cin = x_nhwc.shape
    if stride == 1:
        Ho, Wo = H, W
    elif pad == 1:
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    else:
        Ho, Wo = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1
    out = torch.empty((B * Ho * Wo, cout), device=x_nhwc.device, dtype=torch.float16)
    e = _epilogue(out, bias, residual, act, rowgroup_bias, rows_per_group)
    pre = None
    if gn_groups > 0:
        parts = load().gyre_b200_conv3x3_gn_parts(B, H, W, cout, stride, pad, gn_groups)
        if parts > 0:
            # gn_guard (tests): one more sample's worth of NaNs behind the buffer, returned with it - nothing may write there
            pre = torch.full((B + (1 if gn_guard else 0), parts, gn_groups, 2), float("nan"), device=x_nhwc.device,
                             dtype=torch.float32)
            e.gn_out, e.gn_groups, e.gn_nparts = ptr(pre), gn_groups, parts
    check(load().gyre_b200_conv3x3(ptr(x_nhwc), Cin, B, H, W, Cin, ptr(wp), cout, stride, pad, C.byref(e),
                                   stream_ptr(x_nhwc.device)), "conv3x3")
    if gn_groups > 0:
        return out.view(B, Ho, Wo, cout), pre
    return out.view(B, Ho, Wo, cout)


def pack_upconv3x3(w):
    require_cuda(w)
    cout, cin = w.shape[0], w.shape[1]
    n = load().gyre_b200_upconv3x3_packed_elems(cin, cout)
    wp = torch.empty((n,), device=w.device, dtype=torch.float16)
    w = w.contiguous()
    check(load().gyre_b200_pack_upconv3x3(ptr(w), dtype_code(w), cin, cout, ptr(wp), stream_ptr(w.device)),
          "pack_upconv3x3")
    return wp


def upconv2x(x_nhwc, wp4, cout, bias=None):
    """conv3x3(nearest_upsample_2x(x)): x [B, H, W, Cin] fp16 NHWC -> [B, 2H, 2W, Cout] fp16 NHWC."""
    require_cuda(x_nhwc, wp4)
    B, H, W, Cin = x_nhwc.shape
    out = torch.empty((B * 4 * H * W, cout), device=x_nhwc.device, dtype=torch.float16)
    e = _epilogue(out, bias)
    check(load().gyre_b200_upconv2x(ptr(x_nhwc), Cin, B, H, W, Cin, ptr(wp4), cout, C.byref(e),
                                    stream_ptr(x_nhwc.device)), "upconv2x")
    return out.view(B, 2 * H, 2 * W, cout)


def groupnorm(x1, gamma, beta, groups, eps, silu, x2=None):
    """x1 [B, HW, C1] (+ x2 [B, HW, C2]) fp16 -> [B, HW, C1+C2] fp16."""
    require_cuda(x1, gamma, beta)
    B, HW, C1 = x1.shape
    C2 = x2.shape[2] if x2 is not None else 0
    out = torch.empty((B, HW, C1 + C2), device=x1.device, dtype=torch.float16)
    scratch = torch.empty((load().gyre_b200_groupnorm_scratch_floats(B, HW, groups),), device=x1.device,
                          dtype=torch.float32)
    check(load().gyre_b200_groupnorm(ptr(x1), C1, ptr(x2), C2, B, HW, groups, eps, ptr(gamma), ptr(beta),
                                     1 if silu else 0, ptr(out), ptr(scratch), stream_ptr(x1.device)), "groupnorm")
    return out


def groupnorm_pre(x, gamma, beta, groups, eps, silu, pre):
    """GroupNorm of x [B, HW, C] fp16 from the partials `pre` [B, nparts, G, 2] its producing conv3x3 left."""
    require_cuda(x, gamma, beta, pre)
    B, HW, Cc = x.shape
    out = torch.empty_like(x)
    stats = torch.empty((B, groups, 2), device=x.device, dtype=torch.float32)
    check(load().gyre_b200_groupnorm_pre(ptr(x), Cc, B, HW, groups, eps, ptr(gamma), ptr(beta), 1 if silu else 0,
                                         ptr(out), ptr(pre), pre.shape[1], ptr(stats), stream_ptr(x.device)),
          "groupnorm_pre")
    return out


def layernorm(x, gamma, beta, eps=1e-5):
    require_cuda(x, gamma, beta)
    rows, Cc = x.shape
    out = torch.empty_like(x)
    check(load().gyre_b200_layernorm(ptr(x), rows, Cc, eps, ptr(gamma), ptr(beta), ptr(out), stream_ptr(x.device)),
          "layernorm")
    return out


def attention(q, k, v, heads, scale=None):
    """q [B, Nq, C(+)] / k, v [B, Nk, C(+)] fp16 token-major views (last-dim stride 1) -> [B, Nq, C]."""
    require_cuda(q, k, v)
    B, Nq, _ = q.shape
    Nk = k.shape[1]
    Cc = v.shape[2]
    d = Cc // heads
    if scale is None:
        scale = d ** -0.5
    out = torch.empty((B, Nq, Cc), device=q.device, dtype=torch.float16)
    for t in (q, k, v):
        if t.stride(2) != 1 or t.stride(0) != t.shape[1] * t.stride(1):
            raise NativeError("attention operands must be [B, N, *] views with dense batch stride")
    check(load().gyre_b200_attention(ptr(q), q.stride(1), ptr(k), k.stride(1), ptr(v), v.stride(1), B, heads, Nq, Nk,
                                     d, scale, ptr(out), Cc, stream_ptr(q.device)), "attention")
    return out


def tome_merge_kv(k, v, r, return_plan=False):
    """k, v [B, N, C] fp16 (dense) -> merged [B, N - r, C] pair (reference: bipartite_soft_matching + merge_wavg).
    return_plan: also (node_idx [B, Na], unm_idx [B, Na - r], src_idx [B, r]) as int64 tensors - the plan the kernels
    built (include/gyre_b200.h: gyre_b200_tome_plan_offsets)."""
    require_cuda(k, v)
    B, Nt, Cc = k.shape
    r = min(r, Nt // 2)
    n = C.c_size_t()
    check(load().gyre_b200_tome_workspace_bytes(B, Nt, Cc, C.byref(n)), "tome_workspace_bytes")
    ws = torch.empty((n.value,), device=k.device, dtype=torch.uint8)
    ko = torch.empty((B, Nt - r, Cc), device=k.device, dtype=torch.float16)
    vo = torch.empty_like(ko)
    k, v = k.contiguous(), v.contiguous()
    check(load().gyre_b200_tome_merge_kv(ptr(k), ptr(v), B, Nt, Cc, r, ptr(ko), ptr(vo), ptr(ws), ws.numel(),
                                         stream_ptr(k.device)), "tome_merge_kv")
    if not return_plan:
        return ko, vo
    o = [C.c_size_t(), C.c_size_t(), C.c_size_t()]
    check(load().gyre_b200_tome_plan_offsets(B, Nt, Cc, C.byref(o[0]), C.byref(o[1]), C.byref(o[2])), "tome_plan_offsets")
    Na = (Nt + 1) // 2
    arr = [ws[x.value:x.value + B * Na * 4].view(torch.int32).view(B, Na).long() for x in o]
    return ko, vo, (arr[0], arr[1][:, :Na - r], arr[2][:, :r])


FAMILIES = ("gemm", "conv3x3", "attention", "groupnorm", "layernorm", "softmax", "elementwise", "tome")


def launch_count() -> int:
    return int(load().gyre_b200_launch_count())


def set_tunable(name: str, value: int):
    check(load().gyre_b200_set_tunable(name.encode(), int(value)), "set_tunable")


def get_tunable(name: str) -> int:
    v = C.c_int(0)
    check(load().gyre_b200_get_tunable(name.encode(), C.byref(v)), "get_tunable")
    return v.value


def prof_enable(on: bool):
    load().gyre_b200_prof_enable(1 if on else 0)


def prof_reset():
    load().gyre_b200_prof_reset()


def prof_read(peak_tflops=None, peak_gbs=None) -> dict:
    """{family: {count, ms, flops, bytes[, roofline_ms]}} since the last reset (synchronises the device).  With the two
    peaks, `roofline_ms` = sum over the launches of max(flops / peak_tflops, bytes / peak_gbs)."""
    out = {}
    for i, name in enumerate(FAMILIES):
        c, ms, fl, by = C.c_ulonglong(), C.c_double(), C.c_double(), C.c_double()
        check(load().gyre_b200_prof_read(i, C.byref(c), C.byref(ms), C.byref(fl), C.byref(by)), "prof_read")
        out[name] = {"count": c.value, "ms": ms.value, "flops": fl.value, "bytes": by.value}
        if peak_tflops and peak_gbs:
            ideal = C.c_double()
            check(load().gyre_b200_prof_read_roofline(i, float(peak_tflops), float(peak_gbs), C.byref(ideal)),
                  "prof_read_roofline")
            out[name]["roofline_ms"] = ideal.value
    return out
