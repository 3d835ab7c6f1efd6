"""Per-sample generator RNG contract (reference: gyre/pipeline/randtools.py:39-64 `batched_randn`).

The reference draws ONE `torch.randn((1, *shape[1:]))` per generator per call, on the generator's own
device, and concatenates; this is what makes results independent of how requests are batched
(reference tests/batch_independance.py:16-26).  `torch.randn` (Philox / mt19937 through ATen) is kept on
purpose: bit-identical seeds are part of the parity contract, so the draw itself is not re-implemented.
"""
from __future__ import annotations

import torch


def batched_randn(shape, generators, device, dtype):
    if shape[0] % len(generators) != 0:
        raise ValueError(
            f"shape[0] ({shape[0]}) needs to be a multiple of len(generators) ({len(generators)})"
        )
    draws = [
        torch.randn((1, *shape[1:]), generator=g, device=g.device, dtype=dtype)
        for g in list(generators) * (shape[0] // len(generators))
    ]
    return torch.cat(draws, dim=0).to(device)


def predraw_noise(steps, shape, generators, device, dtype):
    """`steps` successive batched_randn draws, stacked: [steps, *shape].

    Each generator is advanced exactly as `steps` separate calls would advance it (the draws of different
    generators are independent streams), so drawing a run's noise up front is indistinguishable from the
    reference's per-step draws - it only removes a host->device copy from every step."""
    if steps <= 0:
        return torch.empty((0, *shape), device=device, dtype=dtype)
    per_gen = []
    reps = shape[0] // len(generators)
    if shape[0] % len(generators) != 0:
        raise ValueError(
            f"shape[0] ({shape[0]}) needs to be a multiple of len(generators) ({len(generators)})"
        )
    if reps != 1:
        # generators are re-used inside one call: keep the reference's exact interleaving
        return torch.stack([batched_randn(shape, generators, device, dtype) for _ in range(steps)])
    for g in generators:
        per_gen.append(torch.stack([
            torch.randn((1, *shape[1:]), generator=g, device=g.device, dtype=dtype)[0] for _ in range(steps)
        ]))
    out = torch.stack(per_gen, dim=1)          # [steps, B, ...]
    if out.device.type == "cpu" and torch.device(device).type == "cuda":
        out = out.pin_memory()
        return out.to(device, non_blocking=True)
    return out.to(device)
