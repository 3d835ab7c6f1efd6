"""Safety-checker oracle (oracle/safety.py) against the committed fixtures of scripts/make_golden.py:pin_safety - Pillow's own
`Image.resize(BICUBIC)` output and the reference's FlagOnlySafetyChecker (gyre/pipeline/safety_checkers.py) - and the
host-side coefficient tables of the product against the oracle's."""
import os

import numpy as np
import pytest
import torch

from oracle import safety as osf

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "safety.pt"))


def _crc(a):
    return int(np.bitwise_xor.reduce((a.astype(np.int64).ravel() * (np.arange(a.size) % 65521 + 1)) % (1 << 31)))


def test_resize_matches_pillow_fixture(gold):
    for r in gold["resize"]:
        h, w = r["image_hw"]
        img = osf.synthetic_image(h, w)
        nh, nw = osf.resize_output_size(h, w, 224)
        assert (nh, nw) == tuple(r["size"])
        got = osf.pil_resize_bicubic(img, nw, nh)
        assert int(got.astype(np.int64).sum()) == r["resized_sum"] and _crc(got) == r["resized_crc"], (h, w)
        if r["resized"] is not None:
            assert np.array_equal(got, r["resized"].numpy())
        pv = torch.from_numpy(osf.clip_preprocess(img[None])[0]).half()
        assert torch.equal(pv, r["pixel_values_f16"]), (h, w)


def test_resize_against_installed_pillow():
    Image = pytest.importorskip("PIL.Image")
    for (h, w, oh, ow) in [(97, 131, 224, 302), (512, 640, 224, 280), (50, 50, 224, 224), (224, 500, 224, 500)]:
        img = osf.synthetic_image(h, w)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
        assert np.array_equal(osf.pil_resize_bicubic(img, ow, oh), ref)


@pytest.mark.parametrize("sizes", [(512, 224), (64, 224), (96, 336), (768, 336), (100, 605), (37, 224), (1024, 224), (513, 224), (7, 3)])
def test_product_tables_equal_oracle_tables(sizes):
    from gyre_b200.safety_checker import pil_bicubic_tables
    b1, k1, s1 = osf.pil_bicubic_coeffs(*sizes)
    b2, k2, s2 = pil_bicubic_tables(*sizes)
    assert s1 == s2 and np.array_equal(b1, b2) and np.array_equal(k1, k2)
    # Pillow's fixed-point rows sum to 1 << 22 up to the rounding of each tap
    assert np.abs(k2.sum(1) - (1 << 22)).max() <= k2.shape[1]


def test_resize_output_size_rule():
    assert osf.resize_output_size(512, 512) == (224, 224)
    assert osf.resize_output_size(768, 512) == (336, 224)
    assert osf.resize_output_size(512, 768) == (224, 336)
    assert osf.resize_output_size(100, 37) == (605, 224)            # int() truncation of 224 * 100 / 37 = 605.4


def test_vision_oracle_and_flag_loop_match_reference_fixture(gold):
    for name, m in gold["models"].items():
        sd = {k: v.float() for k, v in m["state_dict"].items()}
        P = {k[len("vision_model."):] if k.startswith("vision_model.vision_model.") else k: v for k, v in sd.items()}
        vis = m["vision_config"]
        _, emb = osf.clip_vision_forward(P, m["clip_input"].float(), num_layers=vis["num_hidden_layers"],
                                         num_heads=vis["num_attention_heads"], patch_size=vis["patch_size"],
                                         hidden_act=vis["hidden_act"])
        assert (emb - m["image_embeds"]).abs().max().item() < 5e-5, name
        scores = osf.cosine_scores(emb, P)
        assert (scores - m["scores"]).abs().max().item() < 1e-5, name
        res, flags = osf.flag_only(m["scores"].numpy(), sd["special_care_embeds_weights"], sd["concept_embeds_weights"])
        assert flags == m["flags"] and 0 < sum(flags) < len(flags)
        for r, g in zip(res, m["result"]):
            assert [float(v) for v in r["special_scores"].values()] == g["special_scores"]
            assert [float(v) for v in r["concept_scores"].values()] == g["concept_scores"]
            assert r["bad_concepts"] == g["bad_concepts"]


def test_flag_loop_adjustment_after_special_care_hit():
    # a concept 0.005 under its threshold only fires when a special-care concept fired first (adjustment 0.01)
    scores = np.array([[0.5, 0.0, 0.0] + [0.295] + [0.0] * 16, [0.0, 0.0, 0.0] + [0.295] + [0.0] * 16], np.float32)
    res, flags = osf.flag_only(scores, [0.4, 0.4, 0.4], [0.3] * 17)
    assert flags == [True, False]
    assert res[0]["concept_scores"][0] == 0.005 and res[1]["concept_scores"][0] == -0.005
