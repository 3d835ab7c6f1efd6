"""compute-sanitizer smoke over the conv epilogue that leaves GroupNorm statistics (EPI 6) + groupnorm_pre + the WebP encoder:
`compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_gnfuse.py` (small shapes: the tools slow kernels ~100x)."""
import math
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gyre_b200 import _native as N  # noqa: E402
from gyre_b200.images import to_webp_bytes  # noqa: E402

g = torch.Generator("cpu").manual_seed(0)
for (B, H, W, Cin, Cout, stride, bn) in [(2, 32, 32, 64, 320, 1, 160), (1, 32, 64, 64, 128, 1, 128), (2, 64, 64, 64, 640, 2, 160),
                                       (1, 32, 32, 64, 512, 1, 256)]:
    N.set_tunable("FORCE_BN", bn)      # small problems would pick 64-wide tiles, which cannot carry the statistics
    x = torch.randn(B, H, W, Cin, generator=g).half().cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)).half().cuda()
    parts = N.load().gyre_b200_conv3x3_gn_parts(B, H, W, Cout, stride, 1, 32)
    out, pre = N.conv3x3(x, N.pack_conv3x3(w), Cout, stride=stride, gn_groups=32)
    print((B, H, W, Cin, Cout, stride), "parts", parts, "fused" if pre is not None else "not fused")
    if pre is None:
        continue
    Bo, Ho, Wo, _ = out.shape
    o = out.double().view(B, Ho * Wo, 32, Cout // 32)
    assert torch.allclose(pre.double().sum(1)[..., 0], o.sum((1, 3)), rtol=1e-4, atol=1e-2)
    gamma = torch.ones(Cout, device="cuda")
    beta = torch.zeros(Cout, device="cuda")
    y = N.groupnorm_pre(out.view(B, Ho * Wo, Cout), gamma, beta, 32, 1e-5, True, pre)
    assert torch.isfinite(y).all()
N.set_tunable("FORCE_BN", 0)
rng = np.random.default_rng(3)
img = rng.integers(0, 256, (2, 40, 56, 3), dtype=np.uint8)
assert len(to_webp_bytes(torch.from_numpy(img).cuda())) == 2
torch.cuda.synchronize()
print("ok")
