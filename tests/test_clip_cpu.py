"""CLIP text-encoder oracle against golden vectors produced by the installed transformers CLIPTextModel
(scripts/make_golden.py: pin_clip), and the host-side inventory of the native text encoder."""
import os

import pytest
import torch

from oracle import clip as oclip

GOLD = os.path.join(os.path.dirname(__file__), "golden", "clip.pt")


@pytest.mark.parametrize("act", ["quick_gelu", "gelu"])
def test_clip_oracle_vs_transformers_golden(act):
    g = torch.load(GOLD)[act]
    last, hs = oclip.clip_text_forward(g["state_dict"], g["ids"], num_layers=3, num_heads=4, hidden_act=act)
    assert (last - g["last_hidden_state"]).abs().max().item() < 2e-5
    for a, b in zip(hs, g["hidden_states"]):
        assert (a - b).abs().max().item() < 2e-5
    pen = oclip.alt_layer(g["state_dict"], g["ids"], "penultimate", num_layers=3, num_heads=4, hidden_act=act)
    assert (pen - g["penultimate"]).abs().max().item() < 2e-5
    # integer layer n == final_layer_norm(hidden_states[-n]); n = 1 is the last layer's output, i.e. "final"
    assert torch.allclose(oclip.alt_layer(g["state_dict"], g["ids"], 1, num_layers=3, num_heads=4, hidden_act=act), last)


def test_clip_param_inventory_matches_transformers_names():
    from gyre_b200.text_encoder import ClipTextConfig, clip_param_shapes
    g = torch.load(GOLD)["quick_gelu"]
    shapes = clip_param_shapes(ClipTextConfig.tiny())
    assert set(shapes) == set(g["state_dict"])
    for k, v in g["state_dict"].items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    assert sum(torch.Size(v).numel() for v in clip_param_shapes(ClipTextConfig.clip_l()).values()) == 123060480


def test_clip_needs_cuda():
    from gyre_b200 import _native as N
    from gyre_b200.text_encoder import B200CLIPTextModel, ClipTextConfig
    if not torch.cuda.is_available():
        with pytest.raises(N.NativeError):
            B200CLIPTextModel(ClipTextConfig.tiny())
