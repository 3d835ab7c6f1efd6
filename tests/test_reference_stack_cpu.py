"""Drop-in check of the boundary: the REFERENCE's own wrapper classes - CFGUNetFromDiffusersUNet ->
UNetWithEmbeddings -> (UnetWithExtraChannels) -> CFGUNet_Parallel / Sequential, imported from /root/reference - driven
over an object with B200UNet's exact call signature.  /root/reference exists only in the build container, so these
tests skip elsewhere; what they produced is frozen in tests/golden/wrappers.pt for the portable tests."""
import inspect
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import _vendored  # noqa: E402
from fakes import fake_unet_math  # noqa: E402

pytestmark = pytest.mark.skipif(not _vendored.available(), reason="/root/reference is not mounted here")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _standin_class():
    """A CPU object whose __call__ has B200UNet's signature (checked) and returns B200UNet's output type."""
    from gyre_b200.unet import B200UNet, UNetOutput

    class StandIn:
        def __call__(self, latents, t, *, encoder_hidden_states, down_block_additional_residuals=None,
                     mid_block_additional_residual=None, adapter_states=None, added_cond_kwargs=None, **kwargs):
            assert not kwargs, f"unexpected keyword arguments {sorted(kwargs)}"
            return UNetOutput(sample=fake_unet_math(latents, t, encoder_hidden_states))

    want = inspect.signature(B200UNet.forward)
    got = inspect.signature(StandIn.__call__)
    assert [(p.name, p.kind) for p in want.parameters.values()] == [(p.name, p.kind) for p in got.parameters.values()]
    return StandIn


@pytest.mark.parametrize("dt_name,dt", [("fp32", torch.float32), ("fp16", torch.float16)])
@pytest.mark.parametrize("with_extra", [False, True])
def test_reference_wrappers_accept_the_b200_unet_signature(dt_name, dt, with_extra):
    _, _, rcfg, rcore = _vendored.gyre_pipeline_pure()
    W = torch.load(os.path.join(GOLD, "wrappers.pt"))
    I = W["inputs"]
    B = I["lat"].shape[0]
    unet = rcore.CFGUNetFromDiffusersUNet(_standin_class()())
    kids = rcfg.CFGChildUnets(g=rcore.UNetWithEmbeddings(unet, I["cond"].to(dt), "g"),
                              u=rcore.UNetWithEmbeddings(unet, I["unc"].to(dt), "u"),
                              f=rcore.UNetWithEmbeddings(unet, torch.cat([I["unc"], I["cond"]]).to(dt), "f"))
    if with_extra:
        kids = kids.wrap_all(rcore.UnetWithExtraChannels, I["extra"].to(dt))
    key = f"cfg_{dt_name}_{'extra' if with_extra else 'plain'}"
    for t_name, t in (("tvec", I["t_vec"]), ("tint", I["t_int"])):
        assert torch.equal(rcfg.CFGUNet_Parallel(kids, I["scale"], B)(I["lat"].to(dt), t), W[f"{key}_{t_name}_parallel"])
        assert torch.equal(rcfg.CFGUNet_Sequential(kids, I["scale"], B)(I["lat"].to(dt), t), W[f"{key}_{t_name}_sequential"])


def test_reference_controlnet_and_t2i_keywords_are_accepted():
    """UNetWithControlnet / UNetWithT2I hand `down_block_additional_residuals`, `mid_block_additional_residual` and
    `adapter_states` to the UNet as keywords (core.py:55-64, 213-239): B200UNet.__call__ names them."""
    from gyre_b200.unet import B200UNet
    names = set(inspect.signature(B200UNet.forward).parameters)
    assert {"encoder_hidden_states", "down_block_additional_residuals", "mid_block_additional_residual",
            "adapter_states"} <= names
