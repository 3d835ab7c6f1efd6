"""Attention patcher: the reference's `apply_tome(unet)` followed by `unet.r = int`
(nonfree/tome_patcher.py:14-52; gyre/pipeline/unified_pipeline.py:1580-1584).

In the reference this class-swaps every `BasicTransformerBlock.attn1` to
`ToMeMemoryEfficientCrossAttention`, whose forward merges K and V (not Q) with ONE bipartite plan computed
from K before attention (nonfree/tome_memory_efficient_cross_attention.py:22-76).  In the native UNet the
same thing is a per-block `r` handed to gyre_b200_unet_forward: the merge (tome.cu) runs between the fused
QKV projection and the flash-attention kernel.  `apply_tome` therefore only has to arm the UNet object."""
from __future__ import annotations


def apply_tome(model, trace_source: bool = False, prop_attn: bool = True):
    if trace_source:
        raise NotImplementedError("trace_source is not supported by the native ToMe merge")
    model.r = 0
    model._tome_info = {"r": model.r, "size": None, "source": None, "trace_source": False, "prop_attn": prop_attn,
                        "class_token": False, "distill_token": False}
    return model
