"""Diagnostic: whole-loop CUDA graph replay vs the eager loop on the tiny pipeline - which input goes stale?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gyre_b200 import _native as N  # noqa: E402
from gyre_b200.pipeline import B200Pipeline  # noqa: E402
from gyre_b200.unet import B200UNet  # noqa: E402
from oracle.unet import UNetConfig, synth_params, unet_param_shapes  # noqa: E402


def main():
    cfg = UNetConfig.tiny()
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    unet = B200UNet(cfg).load_state_dict(P)
    pipe = B200Pipeline(unet, None)
    pipe.unet_sample_size_override = 16
    emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=torch.Generator().manual_seed(11)).cuda()
    emb2 = torch.randn(2, 77, cfg.cross_attention_dim, generator=torch.Generator().manual_seed(21)).cuda()
    unc = torch.randn(1, 77, cfg.cross_attention_dim, generator=torch.Generator().manual_seed(12)).expand(2, -1, -1).cuda()

    def run(seed0, e, steps, graph, sampler):
        pipe.use_cuda_graph = graph
        gens = [torch.Generator("cpu").manual_seed(seed0 + i) for i in range(2)]
        out = pipe(e, unc, height=128, width=128, num_inference_steps=steps, guidance_scale=7.5, generator=gens,
                   sampler=sampler, output_type="latent", return_fp32_latents=True).latents
        pipe.use_cuda_graph = False
        return out.clone()

    for tun in ({}, {"PDL": 0}, {"CTX_KV_CACHE": 1, "STREAMK": 0, "MCAST": 0, "PDL": 0}):
        for k, v in tun.items():
            N.set_tunable(k, v)
        for sampler in ("k_euler", "k_euler_ancestral"):
            for steps in (1, 2, 9):
                unet.__dict__.pop("_loop_graphs", None)
                cases = {"same": (100, emb), "seed": (200, emb), "ctx": (100, emb2), "both": (200, emb2)}
                eager = {k: run(s, e, steps, False, sampler) for k, (s, e) in cases.items()}
                first = run(100, emb, steps, True, sampler)
                line = [f"capture-run {(first - eager['same']).abs().max().item():.3g}"]
                for k, (s, e) in cases.items():
                    r = run(s, e, steps, True, sampler)
                    line.append(f"{k}: vs-eager {(r - eager[k]).abs().max().item():.3g} vs-eager[same] "
                                f"{(r - eager['same']).abs().max().item():.3g}")
                print(f"tun={tun} {sampler} steps={steps}: " + " | ".join(line), flush=True)


if __name__ == "__main__":
    main()
