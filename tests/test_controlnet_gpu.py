"""ControlNet encoder on the native kernels against the oracle restatement of gyre/pipeline/controlnet/models.py:420-544,
alone and chained into the native UNet the way gyre/pipeline/unet/core.py:213-239 passes the residuals on."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


@pytest.fixture(scope="module")
def setup():
    from oracle.controlnet import controlnet_param_shapes
    from oracle.unet import UNetConfig, synth_params, unet_param_shapes
    from gyre_b200.controlnet import B200ControlNet
    from gyre_b200.controlnet import controlnet_param_shapes as native_shapes
    from gyre_b200.unet import B200UNet
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = UNetConfig.tiny()
    shapes = controlnet_param_shapes(cfg)
    assert shapes == native_shapes(cfg)           # the product keeps its own inventory (it may not import the oracle)
    PC = synth_params(shapes, seed=77)
    # the zero convolutions are zero-initialised in a fresh ControlNet; trained ones are not - use non-trivial values
    PU = synth_params(unet_param_shapes(cfg), seed=1234)
    cn = B200ControlNet(cfg).load_state_dict(PC)
    unet = B200UNet(cfg).load_state_dict(PU)
    return cfg, PC, PU, cn, unet


@pytest.mark.parametrize("B,hw", [(2, 16), (1, 24)])
def test_controlnet_forward_vs_oracle(setup, B, hw):
    from oracle.controlnet import controlnet_forward
    cfg, PC, PU, cn, unet = setup
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 4, hw, hw, generator=g).half()
    ctx = torch.randn(B, 77, cfg.cross_attention_dim, generator=g).half()
    cond = torch.rand(B, 3, 8 * hw, 8 * hw, generator=g).half()
    t = torch.tensor([801, 21][:B])
    with torch.no_grad():
        ref_down, ref_mid = controlnet_forward(PC, cfg, x.float(), t, ctx.float(), cond.float())
    out = cn(x.cuda(), t.cuda(), encoder_hidden_states=ctx.cuda(), controlnet_cond=cond.cuda())
    assert len(out.down_block_res_samples) == len(ref_down) == cn.num_skips
    errs = []
    for k, (mine, ref) in enumerate(zip(out.down_block_res_samples, ref_down)):
        assert tuple(mine.shape) == tuple(ref.shape), k
        errs.append(rel_err(mine.cpu(), ref))
    errs.append(rel_err(out.mid_block_res_sample.cpu(), ref_mid))
    print(f"controlnet B={B} {hw}x{hw}: rel err per residual {['%.2e' % e for e in errs]}")
    assert max(errs) < 5e-3          # the UNet forward itself measures 1.4e-3 against its oracle
    # conditioning_scale and the tuple return (return_dict=False) of the reference signature
    d2, m2 = cn(x.cuda(), t.cuda(), encoder_hidden_states=ctx.cuda(), controlnet_cond=cond.cuda(), conditioning_scale=0.5,
                return_dict=False)
    assert torch.allclose(m2.float(), out.mid_block_res_sample.float() * 0.5, atol=2e-3)


def test_controlnet_into_unet_vs_oracle(setup):
    """The residuals of the native ControlNet conditioning the native UNet == the oracle ControlNet conditioning the
    oracle UNet; and the condition image matters."""
    from oracle.controlnet import controlnet_forward
    from oracle.unet import unet_forward
    cfg, PC, PU, cn, unet = setup
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 4, 16, 16, generator=g).half()
    ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=g).half()
    cond = torch.rand(2, 3, 128, 128, generator=g).half()
    t = torch.tensor([500, 500])
    with torch.no_grad():
        rd, rm = controlnet_forward(PC, cfg, x.float(), t, ctx.float(), cond.float())
        ref = unet_forward(PU, cfg, x.float(), t, ctx.float(), down_block_additional_residuals=list(rd),
                           mid_block_additional_residual=rm)
        plain = unet_forward(PU, cfg, x.float(), t, ctx.float())
    o = cn(x.cuda(), t.cuda(), encoder_hidden_states=ctx.cuda(), controlnet_cond=cond.cuda())
    out = unet(x.cuda(), t.cuda(), encoder_hidden_states=ctx.cuda(),
               down_block_additional_residuals=list(o.down_block_res_samples),
               mid_block_additional_residual=o.mid_block_res_sample).sample
    err = rel_err(out.cpu(), ref)
    moved = rel_err(ref, plain)
    print(f"controlnet -> unet: rel err {err:.3e}; the ControlNet moves the output by {moved:.3e}")
    assert err < 5e-3 and moved > 5e-2


def test_controlnet_rejects_wrong_use(setup):
    cfg, PC, PU, cn, unet = setup
    x = torch.zeros(1, 4, 16, 16).half().cuda()
    ctx = torch.zeros(1, 77, cfg.cross_attention_dim).half().cuda()
    with pytest.raises(ValueError):
        cn(x, 10, encoder_hidden_states=ctx, controlnet_cond=torch.zeros(1, 3, 64, 64).half().cuda())
    with pytest.raises(NotImplementedError):
        cn(x, 10, encoder_hidden_states=ctx, controlnet_cond=torch.zeros(1, 3, 128, 128).half().cuda(),
           class_labels=torch.zeros(1))
