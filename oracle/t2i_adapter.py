"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's T2I-adapter encoder
(gyre/pipeline/t2i_adapter/adapter.py: Downsample :36-62, ResnetBlock :65-99, Adapter :102-132; configuration defaults of
T2iAdapter_main, t2i_adapter/models.py:80-88: cin 192, channels (320, 640, 1280, 1280), nums_rb 2, ksize 1, sk True,
use_conv False).  adapter.py is pure torch: scripts/make_golden.py imports it as is and PINS this restatement bit for bit
(tests/golden/t2i_adapter.pt).

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module."""
from __future__ import annotations

import torch.nn.functional as F


def adapter_param_shapes(channels=(320, 640, 1280, 1280), nums_rb=2, cin=192, ksize=1, sk=True, use_conv=False) -> dict:
    """nn.Module state-dict names of `Adapter` -> shapes."""
    ks = {"conv_in.weight": (channels[0], cin, 3, 3), "conv_in.bias": (channels[0],)}
    for i in range(len(channels)):
        for j in range(nums_rb):
            idx = i * nums_rb + j
            down = i != 0 and j == 0
            in_c = channels[i - 1] if down else channels[i]
            out_c = channels[i]
            p = f"body.{idx}"
            if in_c != out_c or not sk:
                ks[f"{p}.in_conv.weight"] = (out_c, in_c, ksize, ksize)
                ks[f"{p}.in_conv.bias"] = (out_c,)
            ks[f"{p}.block1.weight"] = (out_c, out_c, 3, 3)
            ks[f"{p}.block1.bias"] = (out_c,)
            ks[f"{p}.block2.weight"] = (out_c, out_c, ksize, ksize)
            ks[f"{p}.block2.bias"] = (out_c,)
            if not sk:
                ks[f"{p}.skep.weight"] = (out_c, in_c, ksize, ksize)
                ks[f"{p}.skep.bias"] = (out_c,)
            if down and use_conv:
                ks[f"{p}.down_opt.op.weight"] = (in_c, in_c, 3, 3)
                ks[f"{p}.down_opt.op.bias"] = (in_c,)
    return ks


def adapter_forward(P, x, channels=(320, 640, 1280, 1280), nums_rb=2, ksize=1, sk=True, use_conv=False):
    """Adapter.forward (:119-132): PixelUnshuffle(8), conv_in, nums_rb ResnetBlocks per level (the first block of every
    level but the first downsamples), one feature map per level."""
    ps = ksize // 2
    x = F.pixel_unshuffle(x, 8)
    x = F.conv2d(x, P["conv_in.weight"], P["conv_in.bias"], padding=1)
    feats = []
    for i in range(len(channels)):
        for j in range(nums_rb):
            p = f"body.{i * nums_rb + j}"
            if i != 0 and j == 0:                                            # ResnetBlock.forward :88-99
                if use_conv:
                    x = F.conv2d(x, P[f"{p}.down_opt.op.weight"], P[f"{p}.down_opt.op.bias"], stride=2, padding=1)
                else:
                    x = F.avg_pool2d(x, kernel_size=2, stride=2)
            if f"{p}.in_conv.weight" in P:
                x = F.conv2d(x, P[f"{p}.in_conv.weight"], P[f"{p}.in_conv.bias"], padding=ps)
            h = F.relu(F.conv2d(x, P[f"{p}.block1.weight"], P[f"{p}.block1.bias"], padding=1))
            h = F.conv2d(h, P[f"{p}.block2.weight"], P[f"{p}.block2.bias"], padding=ps)
            if f"{p}.skep.weight" in P:
                x = h + F.conv2d(x, P[f"{p}.skep.weight"], P[f"{p}.skep.bias"], padding=ps)
            else:
                x = h + x
        feats.append(x)
    return feats
