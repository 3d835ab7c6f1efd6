#!/usr/bin/env python
"""One CFG-doubled SD1.5 UNet forward (batch 16 = 8 images x CFG) + one VAE decode (batch 8) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off`.  Weights are synthetic (architecture only)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gyre_b200.config import UNetConfig, VAEConfig  # noqa: E402
from gyre_b200.unet import B200UNet  # noqa: E402
from gyre_b200.vae import B200VAE  # noqa: E402
from gyre_b200.weights import synth_state_dict, unet_param_shapes, vae_param_shapes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--no-vae", action="store_true")
ap.add_argument("--no-unet", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B = a.batch
if not a.no_unet:
    ucfg = UNetConfig.sd15()
    unet = B200UNet(ucfg, dev).load_state_dict(synth_state_dict(unet_param_shapes(ucfg), 1234, torch.float16, device=dev))
    x = torch.randn(B, 4, 64, 64, device=dev).half()
    x = torch.cat([x, x]).contiguous()          # CFG-parallel layout: [x ; x], what the sampling loop feeds the UNet
    t = torch.full((2 * B,), 500, device=dev, dtype=torch.int64)
    ctx = torch.randn(2 * B, 77, 768, device=dev).half()
    unet.forward_raw(x, t, ctx, cfg_duplicate=True)
if not a.no_vae:
    vcfg = VAEConfig.sd()
    vae = B200VAE(vcfg, dev).load_state_dict(synth_state_dict(vae_param_shapes(vcfg), 4321, torch.float16, device=dev))
    z = torch.randn(B, 4, 64, 64, device=dev).half()
    vae.decode_raw(z, postprocess=True, want_u8=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
if not a.no_unet:
    unet.forward_raw(x, t, ctx, cfg_duplicate=True)
if not a.no_vae:
    vae.decode_raw(z, postprocess=True, want_u8=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
