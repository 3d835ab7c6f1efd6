"""PNG tail at the headline batch (8 decoded 512x512 RGB images): device time of gyre_b200_png_encode, the end-to-end
`to_png_bytes` call (device encode + copy of the finished files), and the reference's host path on this box's CPU
(`.cpu()` + torchvision.io.encode_png per image, gyre/images.py:93-111), one JSON line.
    python scripts/bench_png.py [--batch 8] [--size 512]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    from gyre_b200.images import encode_png_u8, to_png_bytes, to_uint8_nhwc
    rng = np.random.default_rng(7)
    y, x = np.mgrid[0:a.size, 0:a.size]
    # photograph-like content: smooth structure + sensor-like noise
    imgs = np.stack([np.stack([127 + 100 * np.sin(x / (20 + 5 * i) + c) * np.cos(y / 31 - c) + rng.normal(0, 3 + i % 4, x.shape)
                               for c in range(3)], 0).clip(0, 255) / 255 for i in range(a.batch)]).astype(np.float32)
    img = torch.from_numpy(imgs).cuda()                         # [B, 3, H, W] in [0, 1], what the VAE tail produces
    u8 = to_uint8_nhwc(img)
    out = {"batch": a.batch, "image": a.size}
    for _ in range(3):
        encode_png_u8(u8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        encode_png_u8(u8)
    e1.record()
    torch.cuda.synchronize()
    out["device_encode_ms"] = round(e0.elapsed_time(e1) / a.iters, 4)
    raw = u8.numel()
    out["device_GBps_of_pixels"] = round(raw / out["device_encode_ms"] / 1e6, 1)
    t0 = time.perf_counter()
    for _ in range(5):
        files = to_png_bytes(img)
    torch.cuda.synchronize()
    out["to_png_bytes_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
    out["bytes_per_image"] = int(np.mean([len(f) for f in files]))
    try:
        import torchvision
        t0 = time.perf_counter()
        host = (img.to("cpu").to(torch.float32) * 255).round().to(torch.uint8)
        ref = [torchvision.io.encode_png(h) for h in host]
        out["reference_host_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
        out["reference_bytes_per_image"] = int(np.mean([r.numel() for r in ref]))
        dec = [torchvision.io.decode_image(torch.frombuffer(bytearray(f), dtype=torch.uint8)) for f in files]
        out["decodes_to_reference_pixels"] = all(torch.equal(d, h) for d, h in zip(dec, host))
    except ImportError:
        pass
    print(json.dumps(out))


if __name__ == "__main__":
    main()
