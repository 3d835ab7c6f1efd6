"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's T2I-adapter encoder
(gyre/pipeline/t2i_adapter/adapter.py: Downsample :36-62, ResnetBlock :65-99, Adapter :102-132; configuration defaults of
T2iAdapter_main, t2i_adapter/models.py:80-88: cin 192, channels (320, 640, 1280, 1280), nums_rb 2, ksize 1, sk True,
use_conv False).  adapter.py is pure torch: scripts/make_golden.py imports it as is and PINS this restatement bit for bit
(tests/golden/t2i_adapter.pt).

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module."""
from __future__ import annotations

import torch
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


def adapter_light_param_shapes(channels=(320, 640, 1280, 1280), nums_rb=4, cin=192) -> dict:
    """nn.Module state-dict names of `Adapter_light` (adapter.py:240-263) -> shapes."""
    ks = {}
    for i, c in enumerate(channels):
        in_c, inter = (cin if i == 0 else channels[i - 1]), c // 4
        ks[f"body.{i}.in_conv.weight"], ks[f"body.{i}.in_conv.bias"] = (inter, in_c, 1, 1), (inter,)
        for j in range(nums_rb):
            for b in ("block1", "block2"):
                ks[f"body.{i}.body.{j}.{b}.weight"], ks[f"body.{i}.body.{j}.{b}.bias"] = (inter, inter, 3, 3), (inter,)
        ks[f"body.{i}.out_conv.weight"], ks[f"body.{i}.out_conv.bias"] = (c, inter, 1, 1), (c,)
    return ks


def adapter_light_forward(P, x, channels=(320, 640, 1280, 1280), nums_rb=4):
    """Adapter_light.forward (adapter.py:254-263) over `extractor` (:217-237) and ResnetBlock_light (:202-214): PixelUnshuffle(8);
    per level [AvgPool2d(2)] -> 1x1 in_conv -> nums_rb x (conv3x3, ReLU, conv3x3, + x) -> 1x1 out_conv, which is both the
    level's feature map and the next level's input.  PINNED against the reference class (scripts/make_golden.py:pin_t2i_adapter)."""
    x = F.pixel_unshuffle(x, 8)
    feats = []
    for i in range(len(channels)):
        p = f"body.{i}"
        if i > 0:
            x = F.avg_pool2d(x, kernel_size=2, stride=2)
        x = F.conv2d(x, P[f"{p}.in_conv.weight"], P[f"{p}.in_conv.bias"])
        for j in range(nums_rb):
            h = F.relu(F.conv2d(x, P[f"{p}.body.{j}.block1.weight"], P[f"{p}.body.{j}.block1.bias"], padding=1))
            x = F.conv2d(h, P[f"{p}.body.{j}.block2.weight"], P[f"{p}.body.{j}.block2.bias"], padding=1) + x
        x = F.conv2d(x, P[f"{p}.out_conv.weight"], P[f"{p}.out_conv.bias"])
        feats.append(x)
    return feats


def style_adapter_forward(P, x, num_head=8, num_token=8):
    """StyleAdapter.forward (adapter.py:186-199) over ResidualAttentionBlock (:153-170; nn.MultiheadAttention written out:
    packed in_proj, softmax(q k^T / sqrt d) v per head over the tokens of one sample, out_proj) and the fp32 LayerNorm
    subclass (:135-142).  PINNED against the reference class (scripts/make_golden.py:pin_t2i_adapter)."""
    B, L, D = x.shape
    T = num_token
    d = D // num_head

    def ln(t, p):
        return F.layer_norm(t.float(), (D,), P[f"{p}.weight"].float(), P[f"{p}.bias"].float(), 1e-5).to(t.dtype)
    style = P["style_embedding"] + torch.zeros((B, T, D), device=x.device)
    h = ln(torch.cat([x, style], dim=1), "ln_pre")
    n_layers = 1 + max(int(k.split(".")[1]) for k in P if k.startswith("transformer_layes."))
    for i in range(n_layers):
        p = f"transformer_layes.{i}"
        qkv = F.linear(ln(h, f"{p}.ln_1"), P[f"{p}.attn.in_proj_weight"], P[f"{p}.attn.in_proj_bias"])
        q, k, v = (t.reshape(B, L + T, num_head, d).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
        a = torch.softmax((q * d ** -0.5) @ k.transpose(-1, -2), dim=-1) @ v
        a = a.permute(0, 2, 1, 3).reshape(B, L + T, D)
        h = h + F.linear(a, P[f"{p}.attn.out_proj.weight"], P[f"{p}.attn.out_proj.bias"])
        m = F.linear(ln(h, f"{p}.ln_2"), P[f"{p}.mlp.c_fc.weight"], P[f"{p}.mlp.c_fc.bias"])
        h = h + F.linear(m * torch.sigmoid(1.702 * m), P[f"{p}.mlp.c_proj.weight"], P[f"{p}.mlp.c_proj.bias"])
    return ln(h[:, -T:, :], "ln_post") @ P["proj"]
