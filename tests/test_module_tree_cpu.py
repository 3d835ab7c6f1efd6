"""The nn.Module surface of the drop-in UNet / VAE objects (gyre_b200/module_tree.py): module names of the held
parameter tree, LoRA hook detection and folding against the run-time arithmetic of gyre's LoraHook
(gyre/pipeline/lora.py:99-160: `output + up(down(input)) * alpha / r * scale`), refusal of unknown hooks."""
import pytest
import torch
from torch import nn

from gyre_b200 import module_tree as MT
from gyre_b200.config import UNetConfig
from gyre_b200.weights import synth_state_dict, unet_param_shapes, vae_param_shapes
from gyre_b200.config import VAEConfig


class LoraHook:
    """Attribute-for-attribute what gyre/pipeline/lora.py:99-112 stores (accelerate itself is not installed here)."""

    def __init__(self, id, up_weight, down_weight, r=4, alpha=None, scale=1.0):
        self.id = id
        self._up_weight = up_weight
        self._down_weight = down_weight
        self._r = r
        self._iscale = alpha / r if alpha else 1.0
        self._scale = scale


class SequentialHook:
    def __init__(self, *hooks):
        self.hooks = list(hooks)


class CloneToGPUHook:
    pass


class SomeForwardPatch:
    pass


def _tree():
    cfg = UNetConfig.tiny()
    P = synth_state_dict(unet_param_shapes(cfg), 7)
    root = MT.build_param_tree(nn.Module(), P)
    return cfg, P, root


def test_param_tree_has_the_diffusers_module_names():
    cfg, P, root = _tree()
    mods = dict(root.named_modules())
    q = mods["down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q"]
    assert isinstance(q, nn.Linear) and q.in_features == 64 and q.out_features == 64 and q.bias is None
    out0 = mods["down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_out.0"]
    assert isinstance(out0, nn.Linear) and out0.bias is not None
    c1 = mods["down_blocks.0.resnets.0.conv1"]
    assert isinstance(c1, nn.Conv2d) and c1.kernel_size == (3, 3) and c1.padding == (1, 1) and c1.stride == (1, 1)
    assert mods["down_blocks.0.downsamplers.0.conv"].stride == (2, 2)
    assert isinstance(mods["down_blocks.0.resnets.0.norm1"], MT.AffineParams)
    # every parameter of the state dict is reachable under its own name, without a copy
    named = dict(root.named_parameters())
    assert set(named) == set(P)
    k = "mid_block.resnets.0.conv1.weight"
    assert named[k].data_ptr() == P[k].data_ptr() and not named[k].requires_grad
    assert sum(p.numel() for p in root.parameters()) == sum(v.numel() for v in P.values())
    # the VAE tree too
    VP = synth_state_dict(vae_param_shapes(VAEConfig.tiny()), 3)
    vroot = MT.build_param_tree(nn.Module(), VP)
    assert set(dict(vroot.named_parameters())) == set(VP)


def test_lora_fold_equals_the_hook_arithmetic():
    cfg, P, root = _tree()
    mods = dict(root.named_modules())
    g = torch.Generator().manual_seed(1)
    lin = mods["down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_k"]
    conv = mods["up_blocks.1.resnets.0.conv1"]
    r = 4
    h_lin = LoraHook(0, torch.randn(lin.out_features, r, generator=g) * 0.1, torch.randn(r, lin.in_features, generator=g) * 0.1,
                     r=r, alpha=2.0, scale=0.7)
    h_conv = LoraHook(0, torch.randn(conv.out_channels, r, 1, 1, generator=g) * 0.1,
                      torch.randn(r, conv.in_channels, 3, 3, generator=g) * 0.1, r=r, alpha=None, scale=1.3)
    h_lin2 = LoraHook(1, torch.randn(lin.out_features, r, generator=g) * 0.1, torch.randn(r, lin.in_features, generator=g) * 0.1,
                      r=r, alpha=8.0, scale=-0.4)
    lin._hf_hook = SequentialHook(CloneToGPUHook(), h_lin, h_lin2)      # two LoRAs on one layer, behind a placement hook
    conv._hf_hook = h_conv
    sig = MT.lora_signature(root)
    assert [e[0] for e in sig] == ["down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_k"] * 2 + ["up_blocks.1.resnets.0.conv1"] or \
        sorted(e[0] for e in sig) == sorted(["down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_k"] * 2 + ["up_blocks.1.resnets.0.conv1"])
    folded = MT.lora_folded_weights(root, sorted({e[0] for e in sig}))
    # Linear: the hook's run-time result
    x = torch.randn(5, lin.in_features, generator=g)
    want = nn.functional.linear(x, lin.weight)
    for h in (h_lin, h_lin2):
        want = want + nn.functional.linear(nn.functional.linear(x, h._down_weight), h._up_weight) * h._iscale * h._scale
    got = nn.functional.linear(x, folded["down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_k.weight"])
    assert torch.allclose(got, want, atol=1e-5)
    # Conv2d: down = conv with the layer's kernel / padding, up = 1x1 conv
    xc = torch.randn(2, conv.in_channels, 6, 6, generator=g)
    wantc = nn.functional.conv2d(xc, conv.weight, padding=1) + \
        nn.functional.conv2d(nn.functional.conv2d(xc, h_conv._down_weight, padding=1), h_conv._up_weight) * h_conv._iscale * h_conv._scale
    gotc = nn.functional.conv2d(xc, folded["up_blocks.1.resnets.0.conv1.weight"], padding=1)
    assert torch.allclose(gotc, wantc, atol=1e-4)
    # a scale change or a removal changes the signature; no hooks -> empty signature
    h_lin.scale = None
    h_lin._scale = 0.1
    assert MT.lora_signature(root) != sig
    del lin._hf_hook, conv._hf_hook
    assert MT.lora_signature(root) == ()


def test_unknown_forward_hooks_are_refused():
    cfg, P, root = _tree()
    m = dict(root.named_modules())["mid_block.attentions.0.proj_in"]
    m._hf_hook = CloneToGPUHook()
    assert MT.lora_signature(root) == ()          # placement hooks are not the native path's business
    m._hf_hook = SomeForwardPatch()
    with pytest.raises(NotImplementedError):
        MT.lora_signature(root)


def test_adopt_module_keeps_names_and_parameters():
    class Orig(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv_in = nn.Conv2d(4, 8, 3, padding=1)
            self.blocks = nn.ModuleList([nn.Linear(8, 8), nn.Linear(8, 8)])

    o = Orig()
    root = MT.adopt_module(nn.Module(), o)
    assert [n for n, _ in root.named_modules()][1:] == [n for n, _ in o.named_modules()][1:]
    assert dict(root.named_parameters())["blocks.1.weight"] is o.blocks[1].weight
