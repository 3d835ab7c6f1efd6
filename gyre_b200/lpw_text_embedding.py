"""Long-prompt-weighting text embedding in front of the native CLIP encoder (reference:
gyre/pipeline/text_embedding/lpw_text_embedding.py:33-403, built by unified_pipeline.py:2269-2304).

Host side (strings and token lists, a few hundred integers per request): the attention-bracket grammar, tokenisation
through the caller's tokenizer, padding to k x 75 + 2 tokens.  Device side: ALL k chunks of all prompts go through
`B200CLIPTextModel.encode` as ONE batch of [B * k, 77] (the reference calls its encoder once per chunk), then one
`gyre_b200_lpw_weight` launch applies the per-token weights and restores each prompt's mean.
"""
from __future__ import annotations

import logging

import torch

from . import _native as N

logger = logging.getLogger(__name__)

ROUND_MULT = 1.1
SQUARE_MULT = 1 / 1.1


def parse_prompt_attention(text: str):
    """`(abc)` x1.1, `(abc:3.12)` x3.12, `[abc]` /1.1, backslash escapes for the bracket characters; unbalanced openers
    apply to the end of the prompt; runs of equal weight are merged.  Returns [[text, weight], ...]
    (lpw_text_embedding.py:33-118 - a hand-written scanner over the same grammar as the reference's regular expression)."""
    res = []
    round_open, square_open = [], []

    def scale_from(start, mult):
        for item in res[start:]:
            item[1] *= mult

    i, n = 0, len(text)
    while i < n:
        ch = text[i]
        if ch == "\\":
            if i + 1 < n and text[i + 1] in "()[]\\":
                res.append([text[i + 1], 1.0])
                i += 2
            else:
                res.append(["", 1.0])          # a lone backslash is dropped (the regex matches it as an empty escape)
                i += 1
            continue
        if ch == "(":
            round_open.append(len(res))
            i += 1
            continue
        if ch == "[":
            square_open.append(len(res))
            i += 1
            continue
        if ch == ":":
            # ":<number>)" closes a round bracket with an explicit weight
            j = i + 1
            if j < n and text[j] in "+-":
                j += 1
            k = j
            while k < n and (text[k].isdigit() and text[k].isascii() or text[k] == "."):
                k += 1
            if k > j and k < n and text[k] == ")":
                weight = text[i + 1:k]
                if round_open:
                    scale_from(round_open.pop(), float(weight))
                else:
                    res.append([text[i:k + 1], 1.0])
                i = k + 1
                continue
            res.append([":", 1.0])
            i += 1
            continue
        if ch == ")":
            if round_open:
                scale_from(round_open.pop(), ROUND_MULT)
            else:
                res.append([")", 1.0])
            i += 1
            continue
        if ch == "]":
            if square_open:
                scale_from(square_open.pop(), SQUARE_MULT)
            else:
                res.append(["]", 1.0])
            i += 1
            continue
        j = i
        while j < n and text[j] not in "\\()[]:":
            j += 1
        res.append([text[i:j], 1.0])
        i = j
    for pos in round_open:
        scale_from(pos, ROUND_MULT)
    for pos in square_open:
        scale_from(pos, SQUARE_MULT)
    if not res:
        res = [["", 1.0]]
    merged = [res[0]]
    for item in res[1:]:
        if item[1] == merged[-1][1]:
            merged[-1][0] += item[0]
        else:
            merged.append(item)
    return merged


def get_prompts_with_weights(tokenizer, prompt, max_length: int):
    """Token ids (no BOS / EOS) and one weight per token for every prompt; a prompt is a string in the bracket grammar
    or an already parsed [(text, weight), ...] list (gyre's `Prompt.as_tokens()`) (:121-158)."""
    tokens, weights, truncated = [], [], False
    for text in prompt:
        pairs = parse_prompt_attention(text) if isinstance(text, str) else text
        ids, ws = [], []
        for word, weight in pairs:
            piece = tokenizer(word).input_ids[1:-1]
            ids += piece
            ws += [weight] * len(piece)
            if len(ids) > max_length:
                truncated = True
                break
        if len(ids) > max_length:
            truncated = True
            ids, ws = ids[:max_length], ws[:max_length]
        tokens.append(ids)
        weights.append(ws)
    if truncated:
        logger.warning("Prompt was truncated. Try to shorten the prompt or increase max_embeddings_multiples")
    return tokens, weights


def pad_tokens_and_weights(tokens, weights, max_length, bos, eos, no_boseos_middle=True, chunk_length=77):
    """BOS + tokens + EOS padding to max_length; weights padded with 1.0 - with BOS / EOS slots inside every chunk unless
    `no_boseos_middle` (:161-192)."""
    multiples = (max_length - 2) // (chunk_length - 2)
    weights_length = max_length if no_boseos_middle else multiples * chunk_length
    out_t, out_w = [], []
    for ids, ws in zip(tokens, weights):
        out_t.append([bos] + ids + [eos] * (max_length - 1 - len(ids)))
        if no_boseos_middle:
            out_w.append([1.0] + ws + [1.0] * (max_length - 1 - len(ws)))
            continue
        if not ws:
            out_w.append([1.0] * weights_length)
            continue
        w = []
        body = chunk_length - 2
        for j in range(multiples):
            w.append(1.0)
            w += ws[j * body:min(len(ws), (j + 1) * body)]
            w.append(1.0)
        w += [1.0] * (weights_length - len(w))
        out_w.append(w)
    return out_t, out_w


def get_unweighted_text_embeddings(text_encoder, text_input, chunk_length: int, no_boseos_middle=True, clip_layer="final"):
    """[B, k * 75 + 2] token ids -> embeddings, k chunks of 77 with BOS / EOS re-inserted at the chunk ends (:195-235).
    All chunks run as one [B * k, 77] batch on the native encoder."""
    B = text_input.shape[0]
    multiples = (text_input.shape[1] - 2) // (chunk_length - 2)
    if multiples <= 1:
        return text_encoder.encode(text_input, clip_layer)
    body = chunk_length - 2
    chunks = []
    for i in range(multiples):
        c = text_input[:, i * body:(i + 1) * body + 2].clone()
        c[:, 0] = text_input[0, 0]
        c[:, -1] = text_input[0, -1]
        chunks.append(c)
    stacked = torch.stack(chunks, dim=1).reshape(B * multiples, chunk_length)
    emb = text_encoder.encode(stacked, clip_layer).reshape(B, multiples, chunk_length, -1)
    if not no_boseos_middle:
        return emb.reshape(B, multiples * chunk_length, -1)
    parts = []
    for i in range(multiples):
        e = emb[:, i]
        if i == 0:
            e = e[:, :-1]
        elif i == multiples - 1:
            e = e[:, 1:]
        else:
            e = e[:, 1:-1]
        parts.append(e)
    return torch.cat(parts, dim=1).contiguous()


def apply_weights(embeddings, weights):
    """`emb *= w; emb *= previous_mean / emb.mean()` per prompt (:352-371) - one native launch."""
    N.require_cuda(embeddings)
    emb = embeddings.to(torch.float16).contiguous()
    B, L, Cc = emb.shape
    w = torch.as_tensor(weights, dtype=torch.float32).to(emb.device).contiguous()
    if tuple(w.shape) != (B, L):
        raise ValueError(f"weights {tuple(w.shape)} do not match the embeddings {(B, L)}")
    out = torch.empty_like(emb)
    N.check(N.load().gyre_b200_lpw_weight(N.ptr(emb), N.ptr(w), B, L, Cc, N.ptr(out), N.stream_ptr(emb.device)), "lpw_weight")
    return out


def get_weighted_text_embeddings(tokenizer, text_encoder, uncond_encoder, device, prompt, uncond_prompt=None,
                                 max_embeddings_multiples=1, no_boseos_middle=False, skip_parsing=False,
                                 skip_weighting=False, clip_layer="final", **kwargs):
    """lpw_text_embedding.py:238-386 with `text_encoder` / `uncond_encoder` = B200CLIPTextModel."""
    max_length = (tokenizer.model_max_length - 2) * max_embeddings_multiples + 2
    if isinstance(prompt, str):
        prompt = [prompt]
    if uncond_prompt is not None and isinstance(uncond_prompt, str):
        uncond_prompt = [uncond_prompt]
    uncond_tokens = uncond_weights = None
    if not skip_parsing:
        prompt_tokens, prompt_weights = get_prompts_with_weights(tokenizer, prompt, max_length - 2)
        if uncond_prompt is not None:
            uncond_tokens, uncond_weights = get_prompts_with_weights(tokenizer, uncond_prompt, max_length - 2)
    else:
        prompt_tokens = [t[1:-1] for t in tokenizer(prompt, max_length=max_length, truncation=True).input_ids]
        prompt_weights = [[1.0] * len(t) for t in prompt_tokens]
        if uncond_prompt is not None:
            uncond_tokens = [t[1:-1] for t in tokenizer(uncond_prompt, max_length=max_length, truncation=True).input_ids]
            uncond_weights = [[1.0] * len(t) for t in uncond_tokens]
    # round the longest prompt up to a multiple of 75 tokens
    longest = max(len(t) for t in prompt_tokens)
    if uncond_prompt is not None:
        longest = max(longest, max(len(t) for t in uncond_tokens))
    body = tokenizer.model_max_length - 2
    max_embeddings_multiples = max(1, min(max_embeddings_multiples, (longest - 1) // body + 1))
    max_length = body * max_embeddings_multiples + 2
    bos, eos = tokenizer.bos_token_id, tokenizer.eos_token_id
    pad = dict(no_boseos_middle=no_boseos_middle, chunk_length=tokenizer.model_max_length)
    prompt_tokens, prompt_weights = pad_tokens_and_weights(prompt_tokens, prompt_weights, max_length, bos, eos, **pad)
    text_embeddings = get_unweighted_text_embeddings(text_encoder, torch.tensor(prompt_tokens, dtype=torch.long, device=device),
                                                     tokenizer.model_max_length, no_boseos_middle, clip_layer)
    uncond_embeddings = None
    if uncond_prompt is not None:
        uncond_tokens, uncond_weights = pad_tokens_and_weights(uncond_tokens, uncond_weights, max_length, bos, eos, **pad)
        uncond_embeddings = get_unweighted_text_embeddings(uncond_encoder,
                                                           torch.tensor(uncond_tokens, dtype=torch.long, device=device),
                                                           tokenizer.model_max_length, no_boseos_middle, clip_layer)
    if (not skip_parsing) and (not skip_weighting):
        text_embeddings = apply_weights(text_embeddings, prompt_weights)
        if uncond_prompt is not None:
            uncond_embeddings = apply_weights(uncond_embeddings, uncond_weights)
    return text_embeddings, uncond_embeddings


class LPWTextEmbedding:
    """`LPWTextEmbedding(max_embeddings_multiples, tokenizer=, text_encoder=, uncond_encoder=, device=)`
    (lpw_text_embedding.py:389-403 over text_embedding.py:1-30)."""

    def __init__(self, max_embeddings_multiples, tokenizer, text_encoder, uncond_encoder=None, device=None,
                 clip_layer="final", **kwargs):
        self.tokenizer = tokenizer
        self.text_encoder = text_encoder
        self.uncond_encoder = uncond_encoder if uncond_encoder is not None else text_encoder
        self.device = device if device is not None else text_encoder.device
        self.max_embeddings_multiples = max_embeddings_multiples
        self.clip_layer = clip_layer

    def get_embeddings(self, prompt, uncond_prompt=None):
        as_tokens = lambda p: p.as_tokens() if hasattr(p, "as_tokens") else p
        return get_weighted_text_embeddings(
            tokenizer=self.tokenizer, text_encoder=self.text_encoder, uncond_encoder=self.uncond_encoder, device=self.device,
            prompt=as_tokens(prompt), uncond_prompt=as_tokens(uncond_prompt) if uncond_prompt is not None else None,
            max_embeddings_multiples=self.max_embeddings_multiples, clip_layer=self.clip_layer)

    def repeat(self, embedding, count):
        bs_embed, seq_len, _ = embedding.shape
        return embedding.repeat(1, count, 1).view(bs_embed * count, seq_len, -1)
