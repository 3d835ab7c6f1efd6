"""Outpaint image tail on the device against the reference's numpy histogram matching (tests/golden/images.pt, produced by
scripts/make_golden.py from gyre/match_histograms.py inside the statements of unified_pipeline.py:2493-2510)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "images.pt")


@pytest.mark.parametrize("name", ["small/fp16", "batch3/fp16"])
def test_outpaint_histogram_match_bit_exact(name):
    from gyre_b200.images import match_histograms_outpaint
    v = torch.load(GOLD)[name]
    out = match_histograms_outpaint(v["result"].cuda(), v["source"].cuda(), v["outmask"].cuda())
    assert out.dtype == torch.float16
    diff = (out.cpu().float() - v["final"].float()).abs()
    assert torch.equal(out.cpu(), v["final"]), f"{int((diff > 0).sum())} of {diff.numel()} pixels differ, max {diff.max().item()}"
    # broadcast inputs ([1, C, H, W] source / mask with more channels, as the pipeline holds them) take the same path
    src4 = torch.cat([v["source"][:1], torch.ones_like(v["source"][:1, :1])], dim=1)
    out2 = match_histograms_outpaint(v["result"].cuda(), src4.cuda(), v["outmask"][:1].cuda())
    assert torch.equal(out2.cpu(), v["final"])


def test_to_uint8_nhwc():
    from gyre_b200.images import to_uint8_nhwc
    x = torch.rand(2, 3, 8, 8, generator=torch.Generator().manual_seed(1)).half()
    ref = (x.permute(0, 2, 3, 1).to(torch.float32) * 255).round().to(torch.uint8)
    assert torch.equal(to_uint8_nhwc(x.cuda()).cpu(), ref)
