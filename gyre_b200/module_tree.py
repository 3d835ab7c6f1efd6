"""The `nn.Module` surface the reference pipeline expects of its UNet / VAE objects ("attributes the
pipeline reads off the UNet object"): `.modules()` / `.named_modules()` (LoRA application and removal,
gyre/pipeline/unified_pipeline.py:2198-2200, gyre/pipeline/lora.py:166-180; reversible-attention context :1457-1469),
`.parameters()` (:122-124), `.dtype` / `.device`, `_hf_hook` look-ups (:1660-1667).

The native models own a PACKED copy of the weights (NHWC / K-major fp16).  The objects here additionally HOLD the
original parameters under the diffusers module names - either the caller's own `nn.Module` tree (`adopt_module`) or a
tree of plain `nn.Linear` / `nn.Conv2d` / affine holders rebuilt from a state dict (`build_param_tree`) - so that

* gyre's LoRA - an accelerate hook per targeted `nn.Linear` / `nn.Conv2d` computing
  `output + up(down(input)) * alpha / r * scale` (gyre/pipeline/lora.py:99-166) - still finds its modules, and
* the native path can see those hooks before every forward and FOLD them, `W += scale * alpha / r * up . down`
  (`lora_folded_weights`), re-packing only the layers whose hooks changed.  A hook that is neither a LoRA hook nor one of
  accelerate's placement hooks changes the computation in a way the native path cannot follow: it raises.
"""
from __future__ import annotations

import torch
from torch import nn

# accelerate / gyre hooks that only move tensors between devices (gyre/pipeline/model_utils.py:200-268)
_PLACEMENT_HOOKS = {"AlignDevicesHook", "CloneToGPUHook", "CpuOffload", "ModelHook", "SequentialHook", "UserCpuOffloadHook"}


class AffineParams(nn.Module):
    """Holder of a norm layer's `weight` / `bias` (GroupNorm / LayerNorm of the original tree)."""

    def forward(self, *a, **k):  # pragma: no cover - the torch forward is never the product path
        raise RuntimeError("gyre_b200 parameter holders are not executable: the forward runs in libgyre_b200")


def _leaf_for(path, weight):
    if weight.ndim == 4:
        out_c, in_c, kh, kw = weight.shape
        stride = 2 if ("downsamplers" in path or (path and path[-1] == "conv" and "down" in ".".join(path))) and kh == 3 else 1
        m = nn.Conv2d(in_c, out_c, (kh, kw), stride=stride, padding=(kh // 2, kw // 2), bias=False, device="meta")
    elif weight.ndim == 2:
        m = nn.Linear(weight.shape[1], weight.shape[0], bias=False, device="meta")
    else:
        m = AffineParams()
    m._parameters.pop("weight", None)
    return m


def build_param_tree(root: nn.Module, state_dict, device=None, dtype=None):
    """Registers every `a.b.c.weight|bias` of a diffusers-layout state dict on `root` as parameters of nested modules
    `root.a.b.c` (4-D weights -> nn.Conv2d, 2-D -> nn.Linear, 1-D -> AffineParams), so that `root.named_modules()`
    yields the diffusers names.  Parameters are frozen (inference only) and are NOT copied when they already live on
    `device` in `dtype`."""
    keys = sorted(state_dict.keys(), key=lambda k: (not k.endswith("weight"), k))
    for key in keys:
        t = state_dict[key]
        if device is not None or dtype is not None:
            t = t.to(device=device, dtype=dtype if t.is_floating_point() else None)
        parts = key.split(".")
        if len(parts) < 2:
            raise ValueError(f"unexpected parameter name {key!r}")
        *path, pname = parts
        node = root
        for i, p in enumerate(path):
            child = node._modules.get(p)
            if child is None:
                last = i == len(path) - 1
                child = _leaf_for(path, t) if (last and pname == "weight") else (AffineParams() if last else nn.Module())
                node.add_module(p, child)
            node = child
        if pname in node._parameters:
            node._parameters.pop(pname)
        node.register_parameter(pname, nn.Parameter(t.detach(), requires_grad=False))
    return root


def adopt_module(root: nn.Module, module: nn.Module):
    """Registers the children of the caller's original model directly on `root`: `root.named_modules()` then lists the
    original names, the parameters stay the caller's tensors (LoRA / textual inversion / clone_model keep working)."""
    for name, child in module.named_children():
        root.add_module(name, child)
    for name, p in module.named_parameters(recurse=False):
        root.register_parameter(name, p)
    return root


def iter_hooks(module):
    hk = getattr(module, "_hf_hook", None)
    if hk is None:
        return
    for h in (getattr(hk, "hooks", None) or [hk]):
        yield h


def _is_lora_hook(h):
    return type(h).__name__ == "LoraHook" or all(hasattr(h, a) for a in ("_up_weight", "_down_weight", "_iscale", "_scale"))


def lora_signature(root: nn.Module):
    """Hashable description of every LoRA hook currently attached below `root` (module name, hook identity, scale,
    weight storage): changes exactly when the folded weights have to be rebuilt.  Raises NotImplementedError for hooks
    that alter the forward in a way the native path cannot reproduce."""
    sig = []
    for name, m in root.named_modules():
        for h in iter_hooks(m):
            if _is_lora_hook(h):
                sig.append((name, id(h), float(h._scale), float(h._iscale), int(h._up_weight.data_ptr()),
                            int(h._down_weight.data_ptr())))
            elif type(h).__name__ not in _PLACEMENT_HOOKS:
                raise NotImplementedError(
                    f"module {name or '<root>'} carries a {type(h).__name__} hook: the native path cannot honour hooks "
                    f"that change a submodule's forward (only LoRA hooks, which it folds into the weights)")
    return tuple(sig)


def lora_delta(module: nn.Module, hook) -> torch.Tensor:
    """`up(down(x)) * iscale * scale` as a weight delta of `module` (gyre/pipeline/lora.py:113-160): Linear
    `up [out, r] @ down [r, in]`; Conv2d `down` is a conv with the module's kernel, `up` a 1x1 conv:
    delta[o, i, kh, kw] = sum_r up[o, r] * down[r, i, kh, kw]."""
    up, down = hook._up_weight.detach(), hook._down_weight.detach()
    w = module.weight
    up, down = up.to(w.device, torch.float32), down.to(w.device, torch.float32)
    if w.ndim == 2:
        d = up.reshape(up.shape[0], -1) @ down.reshape(down.shape[0], -1)
    elif w.ndim == 4:
        d = torch.einsum("or,rikl->oikl", up.reshape(up.shape[0], up.shape[1]), down)
    else:
        raise ValueError(f"cannot fold a LoRA into a {w.ndim}-D weight")
    if d.shape != w.shape:
        d = d.reshape(w.shape)
    return d * (float(hook._iscale) * float(hook._scale))


def lora_folded_weights(root: nn.Module, names):
    """{'<module name>.weight': W + sum of its hooks' deltas (fp32)} for the given module names (hooked now or before)."""
    out = {}
    mods = dict(root.named_modules())
    for name in names:
        m = mods[name]
        w = m.weight.detach().float()
        for h in iter_hooks(m):
            if _is_lora_hook(h):
                w = w + lora_delta(m, h)
        out[name + ".weight"] = w
    return out
