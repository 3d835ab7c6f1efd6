"""Oracle: gyre's ToMe K/V merge restated (TEST ONLY).

Follows the vendored /root/reference/nonfree/ToMe/tome/merge.py:18-97 (bipartite_soft_matching),
:210-224 (merge_wavg) and tome/utils.py:80-105 (parse_r), as used by
nonfree/tome_memory_efficient_cross_attention.py:28-50: no class/distill token, `size=None`
on every call so the weighted average degenerates to a plain mean of the merged group.
"""
from __future__ import annotations

import torch


def parse_r(num_layers: int, r):
    """tome/utils.py:80-105."""
    inflect = 0
    if isinstance(r, list):
        if len(r) < num_layers:
            r = r + [0] * (num_layers - len(r))
        return list(r)
    elif isinstance(r, tuple):
        r, inflect = r
    min_val = int(r * (1.0 - inflect))
    max_val = 2 * r - min_val
    step = (max_val - min_val) / (num_layers - 1)
    return [int(min_val + step * i) for i in range(num_layers)]


def bipartite_soft_matching_plan(metric: torch.Tensor, r: int):
    """merge.py:41-64.  Returns (unm_idx [B,t1-r], src_idx [B,r], dst_idx [B,r], r) or None.
    Even tokens are set A, odd tokens set B; each A token's best B by cosine score; the r
    A tokens with the highest best-score are merged into their B."""
    t = metric.shape[1]
    r = min(r, t // 2)
    if r <= 0:
        return None
    metric = metric / metric.norm(dim=-1, keepdim=True)
    a, b = metric[..., ::2, :], metric[..., 1::2, :]
    scores = a @ b.transpose(-1, -2)
    node_max, node_idx = scores.max(dim=-1)
    edge_idx = node_max.argsort(dim=-1, descending=True)
    unm_idx = edge_idx[..., r:]
    src_idx = edge_idx[..., :r]
    dst_idx = node_idx.gather(dim=-1, index=src_idx)
    return unm_idx, src_idx, dst_idx, r


def merge_mean(plan, x: torch.Tensor) -> torch.Tensor:
    """merge.py:66-80 with mode="sum" on x and on ones, then x/size (merge_wavg :217-223).
    Output order: [unmerged A tokens (in argsort order) ..., all B tokens ...]."""
    unm_idx, src_idx, dst_idx, r = plan
    src, dst = x[..., ::2, :], x[..., 1::2, :]
    n, t1, c = src.shape
    unm = src.gather(dim=-2, index=unm_idx[..., None].expand(n, t1 - r, c))
    srcm = src.gather(dim=-2, index=src_idx[..., None].expand(n, r, c))
    dst_sum = dst.scatter_reduce(-2, dst_idx[..., None].expand(n, r, c), srcm, reduce="sum")
    ones = torch.ones_like(dst[..., :1])
    size = ones.scatter_reduce(-2, dst_idx[..., None], torch.ones_like(srcm[..., :1]), reduce="sum")
    return torch.cat([unm, dst_sum / size], dim=1)
