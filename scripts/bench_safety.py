"""Safety-check tail at production size (CompVis checker: CLIP ViT-L/14, 8 decoded 512x512 images): device time of the
feature extractor and of the checker, and the errors the GPU tests bound, printed as one JSON line.
    python scripts/bench_safety.py [--batch 8] [--size 512]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    from gyre_b200.safety_checker import B200FeatureExtractor, B200SafetyChecker, ClipVisionConfig, safety_checker_param_shapes
    cfg = ClipVisionConfig.vit_l14()
    g = torch.Generator().manual_seed(31)
    sd = {}
    for k, shp in safety_checker_param_shapes(cfg).items():
        if k.endswith("weight") and len(shp) == 1:
            sd[k] = 1 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.05 * torch.randn(shp, generator=g)
        elif k.endswith("embeds"):
            sd[k] = torch.randn(shp, generator=g)
        elif k.endswith("embeds_weights"):
            sd[k] = torch.full(shp, 0.1)
        else:
            sd[k] = torch.randn(shp, generator=g) / (shp[-1] if len(shp) > 1 else 1) ** 0.5 * (0.3 if "embedding" in k else 1.0)
    sc = B200SafetyChecker(cfg).load_state_dict(sd)
    fx = B200FeatureExtractor()
    img = torch.rand(a.batch, 3, a.size, a.size, generator=g).cuda()
    out = {"batch": a.batch, "image": a.size, "model": "CLIP ViT-L/14 (24 layers, 257 tokens)"}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters

    out["feature_extractor_ms"] = round(timed(lambda: fx(img)), 4)
    pv = fx(img).pixel_values
    out["checker_ms"] = round(timed(lambda: sc.scores(pv)), 4)
    out["tail_ms_with_host_read"] = round(timed(lambda: sc(clip_input=fx(img).pixel_values, images=None)), 4)
    # tower FLOPs: per token 2 * (4 C^2 + 2 C F) per layer + attention 4 N C
    C_, F_, N_, L_ = cfg.hidden_size, cfg.intermediate_size, 257, cfg.num_hidden_layers
    flops = a.batch * N_ * L_ * (2 * (4 * C_ * C_ + 2 * C_ * F_) + 4 * N_ * C_)
    out["tower_tflops"] = round(flops / out["checker_ms"] / 1e9, 1)
    # host path of the reference for comparison: PIL resize of the same images on this box's CPU
    try:
        import time
        import numpy as np
        from PIL import Image
        u8 = (img.permute(0, 2, 3, 1) * 255).round().to(torch.uint8)
        t0 = time.perf_counter()
        host = u8.cpu().numpy()
        pil = [Image.fromarray(h).resize((224, 224), resample=Image.BICUBIC) for h in host]
        arr = np.stack([np.asarray(p) for p in pil])
        out["host_pil_resize_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
        dev = fx.resize(u8.contiguous(), 224, 224).cpu().numpy()
        out["resize_bit_exact_vs_pillow"] = bool(np.array_equal(arr, dev))
    except ImportError:
        pass
    gold = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "safety.pt")
    if os.path.exists(gold):
        errs = {}
        for name, m in torch.load(gold)["models"].items():
            c = B200SafetyChecker({"vision_config": m["vision_config"], "projection_dim": m["projection_dim"]})
            c.load_state_dict(m["state_dict"])
            errs[name] = round((c.scores(m["clip_input"].cuda()).cpu() - m["scores"]).abs().max().item(), 6)
        out["fixture_score_max_abs_err"] = errs
    print(json.dumps(out))


if __name__ == "__main__":
    main()
