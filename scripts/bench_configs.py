#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs on ONE B200 (parity-test cases, not bench.py lines):
   C3 SD1.5-inpaint 768x768, 64 Euler-a steps, batch 4 (VAE encode x2 + decode)
   C4 SD2.1-768-v, 50 Euler-a steps, ToMe r = N/2 per block, batch 8 of the 16
   C5 SDXL-base topology 1024x1024, 30 Euler-a steps, batch 8 of the 64
Synthetic weights / embeddings; one warm-up run, one timed run each.  Writes gpurun_out/configs.json."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gyre_b200.config import UNetConfig, VAEConfig
from gyre_b200.pipeline import B200Pipeline
from gyre_b200.unet import B200UNet
from gyre_b200.vae import B200VAE
from gyre_b200.weights import synth_state_dict, unet_param_shapes, vae_param_shapes

dev = torch.device("cuda", 0)
which = sys.argv[1:] or ["c3", "c4", "c5"]
vcfg = VAEConfig.sd()
vae = B200VAE(vcfg, dev).load_state_dict(synth_state_dict(vae_param_shapes(vcfg), 4321, dtype=torch.float16, device=dev))
out = {}


def run(name, ucfg, B, hw, steps, gflop_per_image, **kw):
    unet = B200UNet(ucfg, dev).load_state_dict(synth_state_dict(unet_param_shapes(ucfg), 1234, dtype=torch.float16, device=dev))
    pipe = B200Pipeline(unet, vae)
    if "tome" in kw:
        pipe.set_options({"tome": kw.pop("tome")})
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(B, 77, ucfg.cross_attention_dim, generator=g).half().to(dev)
    unc = torch.randn(B, 77, ucfg.cross_attention_dim, generator=g).half().to(dev)
    ms = []
    for it in range(2):
        gens = [torch.Generator(dev).manual_seed(1000 * it + i) for i in range(B)]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe(emb, unc, height=hw, width=hw, num_inference_steps=steps, guidance_scale=7.5, generator=gens,
             sampler="k_euler_ancestral", output_type="uint8", **kw)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ips = B / (ms[-1] / 1e3)
    out[name] = {"images_per_s": ips, "ms_per_batch": ms[-1], "batch": B, "algorithmic_tflops": ips * gflop_per_image / 1e3}
    print(name, out[name], flush=True)
    del pipe, unet
    torch.cuda.empty_cache()


if "c3" in which:
    img = torch.rand(1, 3, 768, 768, device=dev)
    mask = torch.zeros(1, 1, 768, 768, device=dev)
    mask[:, :, 192:576, 256:640] = 1
    run("C3 SD1.5-inpaint 768^2, 64 Euler-a, batch 4", UNetConfig.sd15_inpaint(), 4, 768, 64, 281363.0, image=img,
        mask_image=mask, strength=1.0)
if "c4" in which:
    run("C4 SD2.1-768-v 768^2, 50 Euler-a, ToMe r=N/2, batch 8", UNetConfig.sd21_v(), 8, 768, 50, 197414.0, tome=4608)
if "c5" in which:
    cfg = UNetConfig.sdxl()
    B = 8
    added = {"text_embeds": torch.randn(B, 1280, device=dev), "time_ids": torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * B, device=dev)}
    run("C5 SDXL-base topology 1024^2, 30 Euler-a, batch 8", cfg, B, 1024, 30, 416142.0, added_cond_kwargs=added)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
