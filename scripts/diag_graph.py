"""Diagnostic: where does a CUDA-graph replay of the UNet forward differ from the eager launch sequence?
Bisects over the tunables (PDL, STREAMK, MCAST, attention variant, K/V cache) on the tiny model:
captures ONE forward with fixed buffers, then replays with (a) the same inputs, (b) a new sample, (c) a new context."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gyre_b200 import _native as N  # noqa: E402
from gyre_b200.unet import B200UNet  # noqa: E402
from oracle.unet import UNetConfig, synth_params, unet_param_shapes  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    cfg = UNetConfig.tiny()
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    unet = B200UNet(cfg).load_state_dict(P)
    g = torch.Generator().manual_seed(3)
    B = 4
    xs = [torch.randn(B, 4, 16, 16, generator=g).half().to(dev) for _ in range(2)]
    ctxs = [torch.randn(B, 77, cfg.cross_attention_dim, generator=g).half().to(dev) for _ in range(2)]
    t = torch.tensor([500] * B, device=dev)

    def eager(x, ctx, bound):
        if bound:
            unet.set_context(ctx)
            return unet.forward_raw(x, t, None).clone()
        unet.set_context(None)
        return unet.forward_raw(x, t, ctx).clone()

    combos = [{}, {"PDL": 0}, {"STREAMK": 0}, {"MCAST": 0}, {"ATT_VARIANT": 198}, {"PDL": 0, "STREAMK": 0, "MCAST": 0}]
    for bound in (True, False):
        for tun in combos:
            saved = {k: N.get_tunable(k) for k in tun}
            for k, v in tun.items():
                N.set_tunable(k, v)
            ref = {(i, j): eager(xs[i], ctxs[j], bound) for i in (0, 1) for j in (0, 1)}
            xin, cin = xs[0].clone(), ctxs[0].clone()
            out = torch.empty_like(ref[(0, 0)])
            if bound:
                unet.set_context(cin)
            unet.forward_raw(xin, t, None if bound else cin, out=out)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                unet.forward_raw(xin, t, None if bound else cin, out=out)
            res = {}
            for (i, j) in [(0, 0), (1, 0), (0, 1), (1, 1), (0, 0)]:
                xin.copy_(xs[i])
                cin.copy_(ctxs[j])
                if bound:
                    unet.set_context(cin)
                gr.replay()
                torch.cuda.synchronize()
                res[(i, j)] = (out - ref[(i, j)]).abs().max().item()
            print(f"bound={bound} tun={tun}: replay-vs-eager max abs diff {res}", flush=True)
            for k, v in saved.items():
                N.set_tunable(k, v)


if __name__ == "__main__":
    main()
