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
