"""Oracle: transformers ``CLIPTextModel`` forward restated functionally (TEST ONLY).

The arithmetic lives in the third-party dependency ``transformers ~= 4.28.1`` (/root/reference/pyproject.toml:21),
called by the reference through gyre/pipeline/text_embedding/text_encoder_alt_layer.py:6-36 and
lpw_text_embedding.py:195-386.  transformers (5.5) IS importable in this container, so this restatement is PINNED:
scripts/make_golden.py runs the real CLIPTextModel on seeded weights / ids and asserts equality (tests/golden/clip.pt).
Weights: flat dict with the transformers state-dict names."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def clip_text_forward(P: dict, ids, *, num_layers, num_heads, hidden_act="quick_gelu", eps=1e-5):
    """Returns (last_hidden_state, hidden_states): hidden_states[k] = output after k layers (k = 0: embeddings)."""
    B, L = ids.shape
    h = P["text_model.embeddings.token_embedding.weight"][ids] + P["text_model.embeddings.position_embedding.weight"][:L][None]
    C = h.shape[-1]
    d = C // num_heads
    mask = torch.full((L, L), float("-inf")).triu(1)
    hs = [h]
    for i in range(num_layers):
        p = f"text_model.encoder.layers.{i}"
        n = F.layer_norm(h, (C,), P[f"{p}.layer_norm1.weight"], P[f"{p}.layer_norm1.bias"], eps)
        q = F.linear(n, P[f"{p}.self_attn.q_proj.weight"], P[f"{p}.self_attn.q_proj.bias"]) * d ** -0.5
        k = F.linear(n, P[f"{p}.self_attn.k_proj.weight"], P[f"{p}.self_attn.k_proj.bias"])
        v = F.linear(n, P[f"{p}.self_attn.v_proj.weight"], P[f"{p}.self_attn.v_proj.bias"])
        sp = lambda t: t.reshape(B, L, num_heads, d).permute(0, 2, 1, 3)
        s = sp(q) @ sp(k).transpose(-1, -2) + mask
        o = (torch.softmax(s, dim=-1) @ sp(v)).permute(0, 2, 1, 3).reshape(B, L, C)
        h = h + F.linear(o, P[f"{p}.self_attn.out_proj.weight"], P[f"{p}.self_attn.out_proj.bias"])
        n = F.layer_norm(h, (C,), P[f"{p}.layer_norm2.weight"], P[f"{p}.layer_norm2.bias"], eps)
        m = F.linear(n, P[f"{p}.mlp.fc1.weight"], P[f"{p}.mlp.fc1.bias"])
        m = m * torch.sigmoid(1.702 * m) if hidden_act == "quick_gelu" else F.gelu(m)
        h = h + F.linear(m, P[f"{p}.mlp.fc2.weight"], P[f"{p}.mlp.fc2.bias"])
        hs.append(h)
    last = F.layer_norm(h, (C,), P["text_model.final_layer_norm.weight"], P["text_model.final_layer_norm.bias"], eps)
    return last, hs


def alt_layer(P, ids, layer="final", **kw):
    """TextEncoderAltLayer.__call__ (text_encoder_alt_layer.py:17-36)."""
    last, hs = clip_text_forward(P, ids, **kw)
    if layer == "final":
        return last
    k = 2 if layer == "penultimate" else int(layer)
    C = last.shape[-1]
    return F.layer_norm(hs[-k], (C,), P["text_model.final_layer_norm.weight"], P["text_model.final_layer_norm.bias"],
                        kw.get("eps", 1e-5))
