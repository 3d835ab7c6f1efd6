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
