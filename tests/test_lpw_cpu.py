"""Host side of the LPW text embedding (bracket grammar, token / weight lists, padding) against the vectors the
reference's own gyre/pipeline/text_embedding/lpw_text_embedding.py produced (tests/golden/lpw.pt, scripts/make_golden.py)."""
import os
import re
import zlib

import torch

from gyre_b200 import lpw_text_embedding as lpw

GOLD = os.path.join(os.path.dirname(__file__), "golden", "lpw.pt")


class ToyTokenizer:
    """scripts/make_golden.py:ToyTokenizer."""
    model_max_length = 77
    bos_token_id = 998
    eos_token_id = 999

    class _Out:
        def __init__(self, ids):
            self.input_ids = ids

    @staticmethod
    def _ids(text):
        return [1 + zlib.crc32(w.encode()) % 990 for w in re.findall(r"[a-z0-9]+|[^\sa-z0-9]", text.lower())]

    def __call__(self, text, max_length=None, truncation=False, **_):
        if isinstance(text, (list, tuple)):
            return self._Out([self(t, max_length, truncation).input_ids for t in text])
        ids = [self.bos_token_id] + self._ids(text) + [self.eos_token_id]
        if truncation and max_length is not None and len(ids) > max_length:
            ids = ids[:max_length - 1] + [self.eos_token_id]
        return self._Out(ids)


PROMPTS = [
    "a (very beautiful:1.3) masterpiece, [dull] colours, ((sharp)) focus",
    "an \\(escaped\\) bracket and a lone : colon (unbalanced",
    "plain prompt without any weighting at all",
    " ".join(f"(word{i}:{1 + (i % 7) / 10:.1f}) filler{i}," for i in range(60)),
    "",
]
NEGATIVE = ["blurry, (low quality:1.4), [[watermark]]", "", "text", "(bad:1.2) " * 50, "ugly"]


def test_parse_prompt_attention_matches_reference():
    g = torch.load(GOLD)["parse"]
    assert len(g["cases"]) >= 20
    for text, ref in zip(g["cases"], g["parsed"]):
        assert lpw.parse_prompt_attention(text) == ref, text
    # the reference's doctest values
    assert lpw.parse_prompt_attention("an (important) word") == [["an ", 1.0], ["important", 1.1], [" word", 1.0]]
    assert lpw.parse_prompt_attention("(unnecessary)(parens)") == [["unnecessaryparens", 1.1]]


def test_tokens_weights_and_padding_match_reference():
    g = torch.load(GOLD)
    tok = ToyTokenizer()
    for mult in (1, 3):
        for name, nomid in (("mid", False), ("nomid", True)):
            v = g[f"lpw/mult{mult}/{name}"]
            max_len = 75 * mult + 2
            t, w = lpw.get_prompts_with_weights(tok, PROMPTS, max_len - 2)
            tn, _ = lpw.get_prompts_with_weights(tok, NEGATIVE, max_len - 2)
            longest = max(max(len(x) for x in t), max(len(x) for x in tn))
            m2 = max(1, min(mult, (longest - 1) // 75 + 1))
            pt, pw = lpw.pad_tokens_and_weights(t, w, 75 * m2 + 2, tok.bos_token_id, tok.eos_token_id, no_boseos_middle=nomid,
                                                chunk_length=77)
            assert torch.equal(torch.tensor(pt), v["tokens"])
            assert torch.equal(torch.tensor(pw), v["weights"])
            # pre-parsed prompts (gyre's Prompt.as_tokens()) take the same path
            pre = [lpw.parse_prompt_attention(p) for p in PROMPTS]
            t2, w2 = lpw.get_prompts_with_weights(tok, pre, max_len - 2)
            assert t2 == t and w2 == w
