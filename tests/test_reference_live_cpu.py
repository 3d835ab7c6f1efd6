"""LIVE pin (only where /root/reference is mounted - skipped elsewhere, the frozen vectors in tests/golden/ cover the rest):
the reference's own gyre/pipeline/unified_pipeline.py is imported (scripts/_vendored.py:gyre_unified_pipeline - absent
third-party packages stood in for by empty classes) and UnifiedPipeline.__call__ runs over the oracle UNet; the oracle's
restatement of the whole request must reproduce it."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "scripts"))
import _vendored  # noqa: E402

pytestmark = pytest.mark.skipif(not _vendored.available(), reason="/root/reference is not mounted here")


def test_unified_pipeline_call_runs_from_the_reference_and_matches_the_oracle():
    import make_golden as mg
    from oracle import hires as ohires
    from oracle import sampling as osamp
    from oracle.unet import OracleUNet, UNetConfig, synth_params, unet_param_shapes
    up = _vendored.gyre_unified_pipeline()
    import inspect
    assert up.__file__.startswith(_vendored.REF) and inspect.getsourcefile(up.UnifiedPipeline) == up.__file__
    _, ksamp, _ = _vendored.k_diffusion()
    cfg = UNetConfig.tiny()
    unet = OracleUNet(cfg, synth_params(unet_param_shapes(cfg), seed=1234))
    g = torch.Generator().manual_seed(5)
    emb = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    unc = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    seeds = [7, 8]
    for (H, W), steps in (((128, 128), 4), ((192, 192), 3)):
        z = mg._reference_call(up, unet=unet, vae=None, unc=unc, emb=emb, sampler_fn=ksamp.sample_euler_ancestral, seeds=seeds,
                               height=H, width=W, num_inference_steps=steps)
        cfgu = osamp.CFGParallel(unet, unc, emb, 7.5)
        with torch.no_grad():
            if H > 128:          # above the native size the reference engages its hires fix by default
                mine = ohires.hires_txt2img_latents(cfgu, batch=2, height=H, width=W, sample_size=16, seeds=seeds, steps=steps,
                                                    oos_fraction=0.6)
            else:
                mine = osamp.txt2img_latents(cfgu, batch=2, in_channels=4, height=H, width=W, sample_size=16, seeds=seeds,
                                             steps=steps, sampler="euler_a")
        ref = 0.18215 * z
        assert (ref - mine).abs().max().item() / mine.abs().max().item() < 2e-6, (H, W)
