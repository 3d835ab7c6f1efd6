"""Multi-process path (N > 1) on the CPU with the gloo backend, world_size 2: weight broadcast, contiguous
sharding of independent images, image gather - the only two collectives the hot path has."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gyre_b200 import dist as gdist
from gyre_b200.config import UNetConfig
from gyre_b200.weights import synth_state_dict, unet_param_shapes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shapes = {k: v for k, v in list(unet_param_shapes(UNetConfig.tiny()).items())[:40]}
        sd = synth_state_dict(shapes, 1234) if rank == 0 else None
        got = gdist.broadcast_state_dict(sd, shapes, 0, "cpu", torch.float32)
        ref = synth_state_dict(shapes, 1234)
        ok_bcast = all(torch.equal(got[k], ref[k]) for k in shapes)
        # 5 images over 2 ranks: uneven contiguous shards
        a, b = gdist.shard_range(5, rank, world)
        imgs = torch.stack([torch.full((4, 4, 3), i, dtype=torch.uint8) for i in range(a, b)])
        out = gdist.gather_images(imgs)
        ok_gather = True
        if rank == 0:
            ok_gather = out.shape == (5, 4, 4, 3) and [int(out[i, 0, 0, 0]) for i in range(5)] == [0, 1, 2, 3, 4]
        else:
            ok_gather = out is None
        # even shards take the single-gather path
        a, b = gdist.shard_range(4, rank, world)
        imgs = torch.stack([torch.full((2, 2, 3), 10 + i, dtype=torch.uint8) for i in range(a, b)])
        out = gdist.gather_images(imgs)
        if rank == 0:
            ok_gather = ok_gather and [int(out[i, 0, 0, 0]) for i in range(4)] == [10, 11, 12, 13]
        q.put((rank, ok_bcast, ok_gather, (a, b)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_b, ok_g, span in sorted(res):
        assert ok_b, f"rank {rank}: broadcast mismatch"
        assert ok_g, f"rank {rank}: gather mismatch"
    assert sorted(r[3] for r in res) == [(0, 2), (2, 4)]


def test_single_process_is_identity():
    assert gdist.world() == (0, 1)
    x = torch.zeros(2, 4, 4, 3, dtype=torch.uint8)
    assert gdist.gather_images(x) is x
    sh = {"a.weight": (4, 4)}
    sd = synth_state_dict(sh, 1)
    out = gdist.broadcast_state_dict(sd, sh, 0, "cpu", torch.float32)
    assert torch.equal(out["a.weight"], sd["a.weight"])
