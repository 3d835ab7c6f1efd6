"""Native CLIP text encoder through the C ABI against the oracle, whose arithmetic is pinned to the
installed transformers CLIPTextModel (tests/golden/clip.pt)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "clip.pt")


@pytest.mark.parametrize("act", ["quick_gelu", "gelu"])
def test_clip_tiny_vs_transformers_golden(act):
    from gyre_b200.text_encoder import B200CLIPTextModel, ClipTextConfig
    g = torch.load(GOLD)[act]
    cfg = ClipTextConfig.tiny()
    cfg.hidden_act = act
    enc = B200CLIPTextModel(cfg).load_state_dict(g["state_dict"])
    out = enc(g["ids"].cuda(), output_hidden_states=True)
    ref = g["last_hidden_state"]
    err = (out.last_hidden_state.float().cpu() - ref).abs().max().item()
    print(f"clip tiny ({act}) last_hidden_state max abs err {err:.3e} (|ref| max {ref.abs().max():.2f})")
    assert err < 1.5e-2
    pen = enc.encode(g["ids"].cuda(), "penultimate")
    assert (pen.float().cpu() - g["penultimate"]).abs().max().item() < 1.5e-2
    for k, (a, b) in enumerate(zip(out.hidden_states, g["hidden_states"])):
        e = (a.float().cpu() - b).abs().max().item()
        assert e < 2e-2 * max(1.0, b.abs().max().item()), f"hidden_states[{k}]: {e}"


def test_clip_l_full_size_vs_oracle():
    """CLIP-L (123 M parameters, 12 layers, 77 tokens) with seeded random weights: native fp16 vs the fp32 oracle."""
    from oracle import clip as oclip
    from oracle.unet import synth_params
    from gyre_b200.text_encoder import B200CLIPTextModel, ClipTextConfig, clip_param_shapes
    cfg = ClipTextConfig.clip_l()
    P = synth_params(clip_param_shapes(cfg), seed=55)
    g = torch.Generator().manual_seed(1)
    P["text_model.embeddings.token_embedding.weight"] = 0.02 * torch.randn(cfg.vocab_size, cfg.hidden_size, generator=g)
    P["text_model.embeddings.position_embedding.weight"] = 0.02 * torch.randn(77, cfg.hidden_size, generator=g)
    ids = torch.randint(0, cfg.vocab_size, (4, 77), generator=g)
    enc = B200CLIPTextModel(cfg).load_state_dict(P)
    out = enc.encode(ids.cuda(), "final").float().cpu()
    ref, _ = oclip.clip_text_forward(P, ids, num_layers=12, num_heads=12)
    err = (out - ref).abs().max().item()
    print(f"CLIP-L last_hidden_state max abs err {err:.3e} (|ref| max {ref.abs().max():.2f})")
    assert err < 3e-2      # output of a LayerNorm: O(1) values, fp16 activations through 12 layers
    # batch independence: a prompt's embedding does not depend on its neighbours
    one = enc.encode(ids[2:3].cuda(), "final").float().cpu()
    assert torch.equal(one, out[2:3])
