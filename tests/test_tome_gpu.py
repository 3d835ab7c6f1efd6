"""ToMe K/V merge (tome.cu) through the C ABI against (a) the golden vectors taken from the reference's
vendored tome/merge.py (tests/golden/tome.pt, fp32) and (b) the oracle restatement run on the same
fp16-rounded inputs; then the full UNet with `unet.r` set, against the oracle forward with tome_r."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def nat():
    from gyre_b200 import _native
    _native.load()
    return _native


def _plan_agreement(got, ref):
    """Fraction of output tokens that match the reference within fp16 rounding."""
    ok = ((got.float() - ref.float()).abs().amax(dim=-1) < 4e-3)
    return ok.float().mean().item()


def test_tome_vs_vendored_golden(nat):
    g = torch.load(os.path.join(GOLD, "tome.pt"))
    for name, rec in g.items():
        if name == "parse_r":
            continue
        k, v, r = rec["k"], rec["v"], rec["r"]
        ko, vo = nat.tome_merge_kv(k.half().cuda(), v.half().cuda(), r)
        assert ko.shape == rec["k_merged"].shape, name
        # the golden run used fp32 inputs; ours rounds k/v to fp16 first: compare token-wise, allowing the odd
        # near-tie to pick a different partner
        ak = _plan_agreement(ko.cpu(), rec["k_merged"])
        av = _plan_agreement(vo.cpu(), rec["v_merged"])
        print(f"tome {name}: tokens matching vendored ToMe  k {ak:.3f}  v {av:.3f}")
        assert ak > 0.9 and av > 0.9, name


@pytest.mark.parametrize("B,N,C,r", [(2, 64, 32, 16), (1, 256, 64, 128), (3, 30, 16, 7), (2, 1024, 320, 512),
                                     (2, 4096, 320, 1000), (1, 9216, 320, 4608)])
def test_tome_vs_oracle_same_inputs(nat, B, N, C, r):
    from oracle import tome as otome
    gen = torch.Generator("cpu").manual_seed(N + r)
    k = torch.randn(B, N, C, generator=gen).half()
    v = torch.randn(B, N, C, generator=gen).half()
    ko, vo = nat.tome_merge_kv(k.cuda(), v.cuda(), r)
    kc, vc = k.cuda().float(), v.cuda().float()
    plan = otome.bipartite_soft_matching_plan(kc, r)
    kr, vr = otome.merge_mean(plan, kc), otome.merge_mean(plan, vc)
    ak, av = _plan_agreement(ko, kr), _plan_agreement(vo, vr)
    assert ko.shape == kr.shape
    assert torch.isfinite(ko).all() and torch.isfinite(vo).all()
    # The kept tokens come out in descending-score order; two near-equal scores may swap places under fp16
    # rounding of the metric, which permutes K and V identically and leaves attention unchanged.  So the
    # functional check is the attention result itself, with the position-wise match reported next to it.
    q = torch.randn(B, 64, C, generator=gen).cuda()
    def attend(kk, vv):
        return torch.softmax(q @ kk.float().transpose(1, 2) * C ** -0.5, dim=-1) @ vv.float()
    err = (attend(ko, vo) - attend(kr, vr)).abs().max().item()
    print(f"tome B{B} N{N} C{C} r{r}: position-wise token match k {ak:.4f} v {av:.4f}; attention max abs diff {err:.2e}")
    # a partial merge (r < N/2) additionally lets a near-tie decide WHICH token crosses the rank-r cut
    assert err < (5e-3 if r >= N // 2 else 3e-2)
    if r >= N // 2:          # everything merged: no ordering freedom left
        assert ak > 0.97 and av > 0.97


def test_unet_tiny_tome_vs_golden():
    from oracle.unet import UNetConfig, synth_params, unet_param_shapes
    from gyre_b200.tome_patcher import apply_tome
    from gyre_b200.unet import B200UNet
    cfg = UNetConfig.tiny()
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    unet = B200UNet(cfg).load_state_dict(P)
    g = torch.load(os.path.join(GOLD, "oracle_tiny.pt"))
    base, tome = g["unet_tiny"], g["unet_tiny_tome"]
    apply_tome(unet)
    unet.r = tome["r"]
    out = unet(base["x"].cuda().half(), base["t"].cuda(), encoder_hidden_states=base["ctx"].cuda().half()).sample
    err = ((out.cpu().float() - tome["eps"]).abs().max() / tome["eps"].abs().max()).item()
    err_nomerge = ((out.cpu().float() - base["eps"]).abs().max() / base["eps"].abs().max()).item()
    print(f"tiny unet with ToMe r={tome['r']}: rel err vs oracle-with-ToMe {err:.3e}; distance to the un-merged "
          f"output {err_nomerge:.3e}")
    assert err < 3e-2
    assert err < err_nomerge      # merging really happened, and the reference's way
    unet.r = 0
    out0 = unet(base["x"].cuda().half(), base["t"].cuda(), encoder_hidden_states=base["ctx"].cuda().half()).sample
    assert ((out0.cpu().float() - base["eps"]).abs().max() / base["eps"].abs().max()).item() < 2e-2
