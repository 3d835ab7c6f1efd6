"""Lossless WebP tail at the headline batch (8 x 512x512 RGB): device time of gyre_b200_webp_encode, `to_webp_bytes` end to end,
and libwebp on this box's CPU through Pillow (the reference uses OpenCV's libwebp: gyre/images.py:125-135), one JSON line."""
import io
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))


def main():
    from gyre_b200.images import encode_webp_u8, to_uint8_nhwc, to_webp_bytes
    rng = np.random.default_rng(7)
    y, x = np.mgrid[0:512, 0:512]
    imgs = np.stack([np.stack([127 + 100 * np.sin(x / (20 + 5 * i) + c) * np.cos(y / 31 - c) + rng.normal(0, 3 + i % 4, x.shape)
                               for c in range(3)], 0).clip(0, 255) / 255 for i in range(8)]).astype(np.float32)
    img = torch.from_numpy(imgs).cuda()
    u8 = to_uint8_nhwc(img)
    out = {"batch": 8, "image": 512}
    for _ in range(3):
        encode_webp_u8(u8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        encode_webp_u8(u8)
    e1.record()
    torch.cuda.synchronize()
    out["device_encode_ms"] = round(e0.elapsed_time(e1) / 20, 4)
    t0 = time.perf_counter()
    for _ in range(5):
        files = to_webp_bytes(img)
    out["to_webp_bytes_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
    out["bytes_per_image"] = int(np.mean([len(f) for f in files]))
    try:
        from PIL import Image
        host = u8.cpu().numpy()
        t0 = time.perf_counter()
        sizes = []
        for h in host:
            b = io.BytesIO()
            Image.fromarray(h).save(b, format="WEBP", lossless=True)
            sizes.append(len(b.getvalue()))
        out["host_libwebp_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
        out["host_libwebp_bytes_per_image"] = int(np.mean(sizes))
        dec = [np.asarray(Image.open(io.BytesIO(f)).convert("RGB")) for f in files]
        out["decodes_to_input_pixels"] = all(np.array_equal(d, h) for d, h in zip(dec, host))
    except ImportError:
        pass
    print(json.dumps(out))


if __name__ == "__main__":
    main()
