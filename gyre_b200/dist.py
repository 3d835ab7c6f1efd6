"""Multi-GPU harness for the hot path.  The reference scales by running one pipeline
replica per `torch.device('cuda', i)` fed from a queue (gyre/manager.py:648-651, 2103) and moves weights to
each GPU through host memory (gyre/pipeline/model_utils.py:249-254).  Here: one process per GPU
(`torchrun`), images of a batch are independent units sharded contiguously across ranks with NO per-step
communication, and exactly two collectives exist:

  * `broadcast_state_dict` - once at load: the weights travel rank 0 -> all ranks as one flat fp16 buffer
    over NVLink (ncclBroadcast), instead of N host->device copies;
  * `gather_images`        - per batch: decoded uint8 images (clamp/scale/pack fused into the VAE's last
    kernel) are gathered on rank 0.

Both work on any torch.distributed backend (`gloo` on CPU for the tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total: int, rank: int, world_size: int):
    """Contiguous split of `total` independent units: ranks [0, total % world) get one extra."""
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def broadcast_state_dict(state_dict, shapes: dict, src: int = 0, device="cpu", dtype=torch.float16):
    """Rank `src` passes its state dict; the others pass None.  Returns the same parameters on every rank
    (as views into one flat buffer of `dtype`).  `shapes` (key -> shape) must be identical on all ranks."""
    rank, ws = world()
    keys = sorted(shapes)
    sizes = []
    for k in keys:
        n = 1
        for s in shapes[k]:
            n *= s
        sizes.append(n)
    total = sum(sizes)
    flat = torch.empty((total,), device=device, dtype=dtype)
    if rank == src:
        if state_dict is None:
            raise ValueError("the source rank must provide the state dict")
        off = 0
        for k, n in zip(keys, sizes):
            flat[off:off + n].copy_(state_dict[k].reshape(-1).to(device=device, dtype=dtype))
            off += n
    if ws > 1:
        dist.broadcast(flat, src=src)
    out = {}
    off = 0
    for k, n in zip(keys, sizes):
        out[k] = flat[off:off + n].view(shapes[k])
        off += n
    return out


def gather_images(images_u8: torch.Tensor, dst: int = 0):
    """[b_local, H, W, 3] uint8 per rank -> [sum b_local, H, W, 3] on rank `dst` (None elsewhere).  Ranks may
    hold different b_local (uneven shards): sizes are exchanged first."""
    rank, ws = world()
    if ws == 1:
        return images_u8
    n_local = torch.tensor([images_u8.shape[0]], device=images_u8.device, dtype=torch.int64)
    counts = [torch.zeros_like(n_local) for _ in range(ws)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    if len(set(counts)) == 1:
        out = torch.empty((ws * counts[0], *images_u8.shape[1:]), device=images_u8.device, dtype=images_u8.dtype) \
            if rank == dst else None
        dist.gather(images_u8.contiguous(), list(out.chunk(ws)) if rank == dst else None, dst=dst)
        return out
    mx = max(counts)
    pad = torch.zeros((mx, *images_u8.shape[1:]), device=images_u8.device, dtype=images_u8.dtype)
    pad[:images_u8.shape[0]] = images_u8
    bufs = [torch.empty_like(pad) for _ in range(ws)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])
