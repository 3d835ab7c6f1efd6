"""Oracle: the device lossless-WebP (VP8L) encoder (gyre_b200/csrc/webp.cu) restated on the CPU, byte for byte (TEST ONLY).

What the reference does is `cv.imencode(".webp", image, [cv.IMWRITE_WEBP_QUALITY, 500])` (> 100 = lossless; gyre/images.py:125-135
toWebpBytes, picked by gyre/services/generate.py:73-76 when the client accepts image/webp).  Lossless, so the contract is the
DECODED image; the pin is a decoder: Pillow's libwebp opens what this module emits and returns the input pixels
(tests/test_webp_cpu.py).  The GPU encoder is compared with this module byte for byte.

Stream (VP8L, "WebP Lossless Bitstream Specification"): RIFF / WEBP / VP8L chunk; 0x2f, 14-bit width - 1, 14-bit height - 1,
alpha flag, version 0; ONE transform - the predictor transform with the whole image in one mode (12: clamp(L + T - TL), the
gradient predictor; block size 2^9 so that the mode image is a handful of identical pixels coded with zero-bit prefix codes);
no colour cache, no meta prefix codes; five prefix codes built from the residuals' histograms (green / red / blue [/ alpha]
as length-limited canonical Huffman codes sent with a flat 4-bit code-length code, unused alphabets as one-symbol simple
codes); then green, red, blue, alpha codes per pixel in scan order, LSB-first.  No LZ77 / colour cache: literal-only."""
from __future__ import annotations

import struct

import numpy as np

from .png import canonical_codes, huffman_lengths

CL_ORDER = (17, 18, 0, 1, 2, 3, 4, 5, 16, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
PRED_MODE = 12
SIZE_BITS = 9


class _Bits:
    def __init__(self):
        self.acc, self.n = 0, 0

    def put(self, value, nbits):
        self.acc |= int(value) << self.n
        self.n += nbits

    def bytes(self):
        return self.acc.to_bytes((self.n + 7) // 8, "little")


def residuals(img: np.ndarray) -> np.ndarray:
    """img [H, W, C] uint8 (C = 3 RGB | 4 RGBA) -> predictor-transform residuals [H, W, 4] as (R, G, B, A) components.
    Top-left pixel: predicted 0xff000000 (ARGB); rest of the first row: L; first column: T; elsewhere clamp(L + T - TL)."""
    H, W, C = img.shape
    px = np.full((H, W, 4), 255, np.int32)
    px[..., :C] = img
    pred = np.zeros_like(px)
    pred[0, 0] = (0, 0, 0, 255)
    pred[0, 1:] = px[0, :-1]
    pred[1:, 0] = px[:-1, 0]
    pred[1:, 1:] = np.clip(px[1:, :-1] + px[:-1, 1:] - px[:-1, :-1], 0, 255)
    return ((px - pred) & 255).astype(np.uint8)


def _put_simple(bw, symbol):
    """A one-symbol prefix code (zero bits per use)."""
    bw.put(1, 1)                   # simple code
    bw.put(0, 1)                   # num_symbols - 1
    if symbol < 2:
        bw.put(0, 1)
        bw.put(symbol, 1)
    else:
        bw.put(1, 1)
        bw.put(symbol, 8)


def _put_code(bw, freq, alphabet):
    """Prefix code for `freq` (256 counts): simple for one or two used symbols, else normal.  Returns (lens, codes)."""
    used = [s for s, f in enumerate(freq) if f]
    lens = [0] * alphabet
    if len(used) == 1:
        _put_simple(bw, used[0])
        return lens, [0] * alphabet
    if len(used) == 2:
        bw.put(1, 1)
        bw.put(1, 1)
        bw.put(1, 1)               # first symbol in 8 bits
        bw.put(used[0], 8)
        bw.put(used[1], 8)
        lens[used[0]] = lens[used[1]] = 1
        codes = [0] * alphabet
        codes[used[1]] = 1
        return lens, codes
    l256 = huffman_lengths(list(freq))
    lens[:256] = l256
    codes = canonical_codes(lens)
    bw.put(0, 1)                   # normal code
    bw.put(19 - 4, 4)              # all 19 code-length-code lengths
    for s in CL_ORDER:
        bw.put(4 if s < 16 else 0, 3)
    bw.put(0, 1)                   # max_symbol = alphabet size: every length is sent
    rev4 = lambda v: int(format(v, "04b")[::-1], 2)
    for l in lens:
        bw.put(rev4(l), 4)
    return lens, codes


def encode_webp(img: np.ndarray) -> bytes:
    H, W, C = img.shape
    if C == 1:
        img = np.repeat(img, 3, axis=2)
        C = 3
    assert C in (3, 4) and 1 <= W <= 16384 and 1 <= H <= 16384
    res = residuals(img)
    bw = _Bits()
    bw.put(0x2F, 8)
    bw.put(W - 1, 14)
    bw.put(H - 1, 14)
    bw.put(1 if C == 4 else 0, 1)
    bw.put(0, 3)
    # predictor transform, one mode for the whole image
    bw.put(1, 1)
    bw.put(0, 2)
    bw.put(SIZE_BITS - 2, 3)
    bw.put(0, 1)                                   # mode image: no colour cache
    _put_simple(bw, PRED_MODE)                     # green = the mode
    _put_simple(bw, 0)                             # red
    _put_simple(bw, 0)                             # blue
    _put_simple(bw, 255)                           # alpha
    _put_simple(bw, 0)                             # distance   (all zero-bit codes: the mode pixels cost nothing)
    bw.put(0, 1)                                   # no more transforms
    bw.put(0, 1)                                   # no colour cache
    bw.put(0, 1)                                   # no meta prefix codes
    chan = {"g": res[..., 1].reshape(-1), "r": res[..., 0].reshape(-1), "b": res[..., 2].reshape(-1), "a": res[..., 3].reshape(-1)}
    tabs = {}
    for name, alphabet in (("g", 280), ("r", 256), ("b", 256), ("a", 256)):
        tabs[name] = _put_code(bw, np.bincount(chan[name], minlength=256).tolist(), alphabet)
    _put_simple(bw, 0)                             # distance
    # pixels (python ints as accumulators: slabs keep the shifts short)
    slab, slabs = _Bits(), []
    g, r, b, a = (chan[k].tolist() for k in "grba")
    (gl, gc), (rl, rc), (bl, bc), (al, ac) = (tabs[k] for k in "grba")
    for i in range(H * W):
        slab.put(gc[g[i]], gl[g[i]])
        slab.put(rc[r[i]], rl[r[i]])
        slab.put(bc[b[i]], bl[b[i]])
        slab.put(ac[a[i]], al[a[i]])
        if slab.n >= 4096:
            slabs.append(slab)
            slab = _Bits()
    slabs.append(slab)
    for sl in slabs:
        bw.put(sl.acc, sl.n)
    data = bw.bytes()
    chunk = b"VP8L" + struct.pack("<I", len(data)) + data + (b"\0" if len(data) & 1 else b"")
    return b"RIFF" + struct.pack("<I", 4 + len(chunk)) + b"WEBP" + chunk
