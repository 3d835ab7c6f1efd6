"""Image-space tail of the pipeline on the device (reference: gyre/pipeline/unified_pipeline.py:2491-2531 and
gyre/images.py).  What the reference does after the VAE decode on the host - `.cpu()`, numpy, per-channel histogram
matching for outpaint requests - stays on the GPU here; the only device -> host copy left is the final uint8 image."""
from __future__ import annotations

import torch

from . import _native as N


def match_histograms_outpaint(result_image, source, outmask):
    """unified_pipeline.py:2493-2510: `reference = source * (1 - outmask) + result * outmask`;
    `result = images.match_histograms(result, reference)` (gyre/images.py:667-672 -> gyre/match_histograms.py:12-37: uint8
    quantisation, per-channel CDF match over the whole batch, float64 table truncated to uint8);
    `result = source * (1 - outmask) + result * outmask`.  All tensors [B, 3, H, W] in [0, 1]; three launches, bit-identical
    to the reference's fp16 evaluation."""
    N.require_cuda(result_image, source, outmask)
    B, Cc, H, W = result_image.shape
    if Cc != 3:
        raise ValueError(f"match_histograms_outpaint: expected RGB images, got {Cc} channels")
    res = result_image.to(torch.float16).contiguous()
    src = source[:, [0, 1, 2]].to(device=res.device, dtype=torch.float16).expand(B, -1, -1, -1).contiguous()
    msk = outmask[:, [0, 1, 2]].to(device=res.device, dtype=torch.float16).expand(B, -1, -1, -1).contiguous()
    if tuple(src.shape) != tuple(res.shape) or tuple(msk.shape) != tuple(res.shape):
        raise ValueError("match_histograms_outpaint: source / outmask must match the result image")
    lib = N.load()
    scratch = torch.empty((lib.gyre_b200_outpaint_scratch_bytes(),), device=res.device, dtype=torch.uint8)
    out = torch.empty_like(res)
    N.check(lib.gyre_b200_outpaint_match_histograms(N.ptr(res), N.ptr(src), N.ptr(msk), B, H * W, N.ptr(out), N.ptr(scratch),
                                                    N.stream_ptr(res.device)), "outpaint_match_histograms")
    return out


def to_uint8_nhwc(images):
    """`(x.to(float32) * 255).round().to(uint8)` in BHWC order (gyre/images.py toPIL / toCV / toPngBytes quantisation) on
    the device: what the PNG / WebP encoders consume, 4x smaller than the fp32 tensor the reference copies to the host."""
    return (images.permute(0, 2, 3, 1).to(torch.float32) * 255).round().to(torch.uint8).contiguous()


def encode_png_u8(u8_nhwc):
    """uint8 [B, H, W, C] on the device (C = 1 grey, 2 grey + alpha, 3 RGB, 4 RGBA) -> (files [B, stride] uint8 on the device,
    lengths [B] int64 on the device): one complete PNG file per row (gyre_b200_png_encode)."""
    import ctypes as C
    N.require_cuda(u8_nhwc)
    if u8_nhwc.dtype != torch.uint8 or u8_nhwc.ndim != 4:
        raise ValueError(f"encode_png_u8: want uint8 [B, H, W, C], got {u8_nhwc.dtype} {tuple(u8_nhwc.shape)}")
    x = u8_nhwc.contiguous()
    B, H, W, Cc = x.shape
    lib = N.load()
    ws_bytes, stride = C.c_size_t(), C.c_size_t()
    N.check(lib.gyre_b200_png_sizes(B, H, W, Cc, C.byref(ws_bytes), C.byref(stride)), "png_sizes")
    ws = torch.empty((ws_bytes.value,), device=x.device, dtype=torch.uint8)
    out = torch.empty((B, stride.value), device=x.device, dtype=torch.uint8)
    lengths = torch.empty((B,), device=x.device, dtype=torch.int64)
    with torch.cuda.device(x.device):
        N.check(lib.gyre_b200_png_encode(N.ptr(x), B, H, W, Cc, N.ptr(out), stride.value, N.ptr(lengths), N.ptr(ws),
                                         ws.numel(), N.stream_ptr(x.device)), "png_encode")
    return out, lengths


def to_png_bytes(tensor):
    """gyre/images.py:93-111 toPngBytes: [B, C, H, W] (or [C, H, W]) float images in [0, 1] -> list of PNG files as bytes,
    `(x.to(float32) * 255).round().to(uint8)` like the reference.  The reference copies the float image to the host and runs
    libpng per image; here quantisation, filtering, entropy coding and checksums run on the device and only the finished
    files (about half the size of the uint8 image, an eighth of the fp32 one) cross to the host.  uint8 [B, H, W, C] input is
    taken as is.  Lossless: decodes to the same pixels as the reference's files."""
    if tensor.dtype == torch.uint8:
        u8 = tensor if tensor.ndim == 4 else tensor[None]
    else:
        t = tensor if tensor.ndim == 4 else tensor[None]
        if t.shape[1] not in (1, 3, 4):
            print(f"Don't know how to save PNGs with {t.shape[1]} channels")       # the reference's behaviour (images.py:109-111)
            return []
        u8 = to_uint8_nhwc(t)
    if not u8.is_cuda:
        if not torch.cuda.is_available():
            raise N.NativeError("to_png_bytes needs a CUDA device: there is no CPU path")
        u8 = u8.cuda()
    files, lengths = encode_png_u8(u8)
    lens = lengths.cpu().tolist()
    host = files[:, :max(lens)].cpu().numpy()
    return [host[i, :n].tobytes() for i, n in enumerate(lens)]


def add_text_chunk_to_png_bytes(binary: bytes, key: str, text: str) -> bytes:
    """gyre/images.py:165-183 addTextChunkToPngBytes: a tEXt chunk in front of IEND (host bytes in, host bytes out)."""
    import struct
    import zlib
    body = key.encode("utf-8") + b"\0" + text.encode("utf-8")
    at = binary.rindex(b"IEND") - 4
    chunk = struct.pack(">I", len(body)) + b"tEXt" + body + struct.pack(">I", zlib.crc32(b"tEXt" + body))
    return binary[:at] + chunk + binary[at:]


_RESIZE_TAPS = {}


def _lanczos3_taps(in_sz: int, scale: float, antialiasing: bool, device):
    """ResizeRight's 1-D plan (gyre/src/ResizeRight/resize_right.py:70-118, 128-213) for interp_methods.lanczos3 (support 6)
    with reflect padding folded into the source indices; antialiasing stretches the window by 1 / scale when downscaling.
    fp32 throughout, in ResizeRight's order of operations.  Returns (idx int32 [out, K], w fp32 [out, K], out_sz)."""
    import math
    key = (in_sz, float(scale), bool(antialiasing), str(device))
    hit = _RESIZE_TAPS.get(key)
    if hit is not None:
        return hit
    eps = torch.finfo(torch.float32).eps
    scale = float(scale)
    out_sz = math.ceil(scale * in_sz)

    def lanczos3(x):
        return ((torch.sin(math.pi * x) * torch.sin(math.pi * x / 3) + eps) / ((math.pi ** 2 * x ** 2 / 3) + eps)) * (abs(x) < 3).to(x.dtype)
    grid = torch.arange(out_sz) / scale + (in_sz - 1) / 2 - (out_sz - 1) / (2 * scale)
    support = 6.0
    if antialiasing and scale < 1.0:
        interp, support = (lambda arg: scale * lanczos3(scale * arg)), support / scale
    else:
        interp = lanczos3
    left = (grid - support / 2 - eps).ceil().long()
    fov = left[:, None] + torch.arange(math.ceil(support - eps))
    pad0 = -fov[0, 0].item()
    w = interp((grid + pad0)[:, None] - (fov + pad0))
    tot = w.sum(1, keepdim=True)
    tot[tot == 0] = 1
    w = (w / tot).to(torch.float32)
    idx = torch.where(fov < 0, -fov, fov)
    idx = torch.where(idx > in_sz - 1, 2 * (in_sz - 1) - idx, idx)
    if idx.min() < 0 or idx.max() > in_sz - 1:
        raise ValueError(f"resize: the filter window ({w.shape[1]} taps) needs more reflect padding than a dimension of {in_sz} "
                         f"allows (the reference fails here too)")
    hit = (idx.to(torch.int32).contiguous().to(device), w.contiguous().to(device), out_sz)
    if len(_RESIZE_TAPS) > 128:
        _RESIZE_TAPS.clear()
    _RESIZE_TAPS[key] = hit
    return hit


def resize(tensor, factors, sharpness=1):
    """gyre/images.py:324-340 `resize`: [B, C, H, W] scaled by `factors` (a number or (fy, fx)) with lanczos3 through ResizeRight,
    reflect padding, antialiasing for sharpness 1 (none for 2), clamped to [0, 1], cast back to the input dtype.  One launch per
    resized dimension (gyre_b200_resample_f32), dimensions in ascending scale order like ResizeRight.  sharpness 0 (area
    downscale) is not built."""
    import ctypes as C  # noqa: F401
    if sharpness not in (1, 2):
        raise NotImplementedError("images.resize: sharpness 0 (area downscale) is not built")
    N.require_cuda(tensor)
    if not isinstance(factors, (tuple, list)):
        factors = (factors, factors)
    elif len(factors) == 1:
        factors = (factors[0], factors[0])
    if tensor.ndim != 4:
        raise ValueError(f"images.resize: want [B, C, H, W], got {tuple(tensor.shape)}")
    lib = N.load()
    cur = tensor.to(torch.float32).contiguous()
    st = N.stream_ptr(cur.device)
    for d in sorted((2, 3), key=lambda a: float(factors[a - 2])):
        f = float(factors[d - 2])
        if f == 1.0:
            continue
        B, Cc, H, W = cur.shape
        idx, w, out_sz = _lanczos3_taps(cur.shape[d], f, sharpness == 1, cur.device)
        if d == 2:
            dst = torch.empty((B, Cc, out_sz, W), device=cur.device, dtype=torch.float32)
            n_outer, in_sz, inner = B * Cc, H, W
        else:
            dst = torch.empty((B, Cc, H, out_sz), device=cur.device, dtype=torch.float32)
            n_outer, in_sz, inner = B * Cc * H, W, 1
        with torch.cuda.device(cur.device):
            N.check(lib.gyre_b200_resample_f32(N.ptr(cur), n_outer, in_sz, inner, N.ptr(idx), N.ptr(w), w.shape[1], out_sz, 0,
                                               N.ptr(dst), st), "resample_f32")
        cur = dst
    # (the reference casts back to the input dtype first and clamps after: an fp16 input is clamped in fp16)
    return cur.to(tensor.dtype).clamp(0, 1)


def rescale(tensor, height, width=None, fit="cover", pad_mode="constant", sharpness=1):
    """gyre/images.py:369-408 `rescale`: resize so that the target box is covered ("cover"), contained ("contain") or hit
    exactly ("strict"), then centre-crop the excess / pad what is missing."""
    if width is None:
        width = height
    orig_h, orig_w = tensor.shape[-2], tensor.shape[-1]
    scale_h, scale_w = height / orig_h, width / orig_w
    if fit == "cover":
        scale_h = scale_w = max(scale_h, scale_w)
    elif fit == "contain":
        scale_h = scale_w = min(scale_h, scale_w)
    tensor = resize(tensor, (scale_h, scale_w), sharpness=sharpness)
    res_h, res_w = tensor.shape[-2], tensor.shape[-1]
    err_h, err_w = (height - res_h) // 2, (width - res_w) // 2
    top, left = (-err_h if err_h < 0 else 0), (-err_w if err_w < 0 else 0)
    tensor = tensor[:, :, top:top + height, left:left + width]
    pad = [err_w, width - res_w - err_w] if err_w > 0 else [0, 0]
    pad += [err_h, height - res_h - err_h] if err_h > 0 else [0, 0]
    return torch.nn.functional.pad(tensor, pad, pad_mode)


def encode_webp_u8(u8_nhwc):
    """uint8 [B, H, W, 3 | 4] on the device -> (files [B, stride] uint8, lengths [B] int64): one lossless WebP per row."""
    import ctypes as C
    N.require_cuda(u8_nhwc)
    if u8_nhwc.dtype != torch.uint8 or u8_nhwc.ndim != 4:
        raise ValueError(f"encode_webp_u8: want uint8 [B, H, W, C], got {u8_nhwc.dtype} {tuple(u8_nhwc.shape)}")
    x = u8_nhwc.contiguous()
    B, H, W, Cc = x.shape
    lib = N.load()
    ws_bytes, stride = C.c_size_t(), C.c_size_t()
    N.check(lib.gyre_b200_webp_sizes(B, H, W, Cc, C.byref(ws_bytes), C.byref(stride)), "webp_sizes")
    ws = torch.empty((ws_bytes.value,), device=x.device, dtype=torch.uint8)
    out = torch.empty((B, stride.value), device=x.device, dtype=torch.uint8)
    lengths = torch.empty((B,), device=x.device, dtype=torch.int64)
    with torch.cuda.device(x.device):
        N.check(lib.gyre_b200_webp_encode(N.ptr(x), B, H, W, Cc, N.ptr(out), stride.value, N.ptr(lengths), N.ptr(ws),
                                          ws.numel(), N.stream_ptr(x.device)), "webp_encode")
    return out, lengths


def to_webp_bytes(tensor):
    """gyre/images.py:125-135 toWebpBytes (lossless): [B, C, H, W] / [C, H, W] float images in [0, 1], quantised
    `(x.to(float32) * 255).round()` like toCV, 1 / 3 / 4 channels (grey is written as RGB) -> list of WebP files as bytes,
    encoded on the device.  uint8 [B, H, W, C] input is taken as is."""
    if tensor.dtype == torch.uint8:
        u8 = tensor if tensor.ndim == 4 else tensor[None]
    else:
        t = tensor if tensor.ndim == 4 else tensor[None]
        u8 = to_uint8_nhwc(t)
    if u8.shape[-1] == 1:
        u8 = u8.expand(-1, -1, -1, 3)
    if not u8.is_cuda:
        if not torch.cuda.is_available():
            raise N.NativeError("to_webp_bytes needs a CUDA device: there is no CPU path")
        u8 = u8.cuda()
    files, lengths = encode_webp_u8(u8)
    lens = lengths.cpu().tolist()
    host = files[:, :max(lens)].cpu().numpy()
    return [host[i, :n].tobytes() for i, n in enumerate(lens)]


def add_text_chunk_to_webp_bytes(binary: bytes, info_fourcc: bytes, text: str) -> bytes:
    """gyre/images.py:186-226 addTextChunkToWebpBytes: a simple lossless file becomes an extended one (VP8X header) with a
    LIST / INFO chunk behind the image (host bytes in, host bytes out)."""
    import struct
    assert len(info_fourcc) == 4, f"{info_fourcc} needs to a be four byte FourCC"
    body = text.encode("utf-8")
    txt_chunk = (b"LIST" + struct.pack("<I", 4 + 4 + 4 + len(body)) + b"INFO" + info_fourcc + struct.pack("<I", len(body)) + body
                 + bytes(len(body) % 2))
    if binary[12:16] != b"VP8L":
        raise NotImplementedError("Can only apply chunk to simple lossless webp")
    remainder = binary[12:]
    info = struct.unpack("<I", binary[21:25])[0]
    width, height, alpha = (info & 0x3FFF) + 1, ((info >> 14) & 0x3FFF) + 1, bool((info >> 28) & 0x1)
    header_chunk = (b"VP8X" + struct.pack("<I", 10) + bytes([0b10000 if alpha else 0, 0, 0, 0]) + struct.pack("<I", width - 1)[0:3]
                    + struct.pack("<I", height - 1)[0:3])
    total = 4 + len(header_chunk) + len(remainder) + len(txt_chunk)
    return b"RIFF" + struct.pack("<I", total) + b"WEBP" + header_chunk + remainder + txt_chunk
