"""Oracle: the device PNG encoder (gyre_b200/csrc/png.cu) restated on the CPU, byte for byte (TEST ONLY).

What the reference does at this point is `torchvision.io.encode_png(image)` (libpng, gyre/images.py:93-111, called per
artifact by gyre/services/generate.py:79).  PNG is lossless, so the contract a replacement has to keep is the DECODED
image, and the pin is the decoders: scripts/make_golden.py:pin_png / tests/test_png_cpu.py check that what this module
emits (a) is accepted by zlib, Pillow and the reference's own decoder `torchvision.io.decode_image` (gyre/images.py:85-90
fromPngBytes) with every CRC verified, and (b) decodes to exactly the pixels `torchvision.io.encode_png` output decodes to.
The GPU encoder is then compared with this module BYTE FOR BYTE.

Stream layout (chosen for one CTA per chunk, no cross-chunk bit dependencies):
  signature, IHDR, one IDAT per chunk of `rows_per_chunk` scanlines, a 4-byte IDAT holding the Adler-32, IEND.
  Chunk = adaptive filter per scanline (minimum sum of absolute residuals, ties to the lower filter type) -> literal-only
  deflate: ONE dynamic-Huffman block (257 literal/length lengths sent with a flat 4-bit code-length code, two 1-bit distance
  codes) or, when that is not smaller, one stored block; then an empty stored block (BFINAL on the last chunk) that
  byte-aligns the stream, like zlib's Z_SYNC_FLUSH.  Chunk 0 starts with the zlib header 78 01."""
from __future__ import annotations

import struct
import zlib

import numpy as np

CHUNK_TARGET = 16384           # filtered bytes per chunk (at least one scanline)
MAX_ROW_BYTES = 32768
MAX_BITS = 15
CL_ORDER = (16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15)
COLOR_TYPE = {1: 0, 2: 4, 3: 2, 4: 6}


def rows_per_chunk(height: int, row_bytes: int) -> int:
    return max(1, min(height, CHUNK_TARGET // row_bytes))


def _filter_rows(img: np.ndarray, r0: int, r1: int) -> np.ndarray:
    """Scanlines r0..r1 of img [H, W*C] (bpp = C bytes) -> filtered bytes with the filter-type byte in front of each."""
    H, wc = img.shape[0], img.shape[1] * img.shape[2]
    bpp = img.shape[2]
    flat = img.reshape(H, wc).astype(np.int32)
    out = np.empty((r1 - r0, 1 + wc), np.uint8)
    for r in range(r0, r1):
        x = flat[r]
        b = flat[r - 1] if r > 0 else np.zeros(wc, np.int32)
        a = np.concatenate([np.zeros(bpp, np.int32), x[:-bpp]]) if wc > bpp else np.zeros(wc, np.int32)
        c = np.concatenate([np.zeros(bpp, np.int32), b[:-bpp]]) if wc > bpp else np.zeros(wc, np.int32)
        p = a + b - c
        pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
        paeth = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))
        cands = [x, x - a, x - b, x - ((a + b) >> 1), x - paeth]
        best, best_cost = 0, None
        for f, res in enumerate(cands):
            v = res & 255
            cost = int(np.where(v < 128, v, 256 - v).sum())
            if best_cost is None or cost < best_cost:
                best, best_cost = f, cost
        out[r - r0, 0] = best
        out[r - r0, 1:] = (cands[best] & 255).astype(np.uint8)
    return out.reshape(-1)


def huffman_lengths(freq) -> list:
    """Code lengths (<= 15) for the symbols with freq > 0, the exact procedure of png.cu: symbols sorted by (freq, symbol),
    two-queue Huffman (on equal weights the LEAF queue is taken first), depths, clamp with the count-shifting repair,
    lengths handed out longest-first to the least frequent."""
    syms = sorted((f, s) for s, f in enumerate(freq) if f > 0)
    n = len(syms)
    lens = [0] * len(freq)
    if n == 1:
        lens[syms[0][1]] = 1
        return lens
    leaf_w = [f for f, _ in syms]
    int_w = []
    par_leaf = [0] * n
    par_int = [0] * (n - 1)
    li = ii = 0
    for k in range(n - 1):
        w = 0
        for _ in range(2):
            if li < n and (ii >= len(int_w) or leaf_w[li] <= int_w[ii]):
                w += leaf_w[li]
                par_leaf[li] = k
                li += 1
            else:
                w += int_w[ii]
                par_int[ii] = k
                ii += 1
        int_w.append(w)
    depth_int = [0] * (n - 1)
    for j in range(n - 3, -1, -1):
        depth_int[j] = depth_int[par_int[j]] + 1
    cnt = [0] * 64
    for i in range(n):
        cnt[min(depth_int[par_leaf[i]] + 1, 63)] += 1
    for d in range(MAX_BITS + 1, 64):
        cnt[MAX_BITS] += cnt[d]
        cnt[d] = 0
    total = sum(cnt[l] << (MAX_BITS - l) for l in range(1, MAX_BITS + 1))
    while total > (1 << MAX_BITS):
        cnt[MAX_BITS] -= 1
        for l in range(MAX_BITS - 1, 0, -1):
            if cnt[l]:
                cnt[l] -= 1
                cnt[l + 1] += 2
                break
        total -= 1
    i = 0
    for l in range(MAX_BITS, 0, -1):
        for _ in range(cnt[l]):
            lens[syms[i][1]] = l
            i += 1
    return lens


def canonical_codes(lens) -> list:
    """Deflate's canonical code (RFC 1951 3.2.2), bit-reversed for LSB-first packing."""
    cnt = [0] * (MAX_BITS + 2)
    for l in lens:
        cnt[l] += 1
    cnt[0] = 0
    nxt = [0] * (MAX_BITS + 2)
    code = 0
    for l in range(1, MAX_BITS + 1):
        code = (code + cnt[l - 1]) << 1
        nxt[l] = code
    out = [0] * len(lens)
    for s, l in enumerate(lens):
        if l:
            c = nxt[l]
            nxt[l] += 1
            out[s] = int(format(c, f"0{l}b")[::-1], 2)
    return out


class _Bits:
    def __init__(self, prefix: bytes = b""):
        self.acc = int.from_bytes(prefix, "little")
        self.n = 8 * len(prefix)

    def put(self, value: int, nbits: int):
        self.acc |= value << self.n
        self.n += nbits

    def align(self):
        self.n = (self.n + 7) & ~7

    def bytes(self) -> bytes:
        return self.acc.to_bytes((self.n + 7) // 8, "little")


def deflate_chunk(data: np.ndarray, first: bool, last: bool) -> bytes:
    freq = np.bincount(data, minlength=257).tolist()
    freq[256] = 1
    lens = huffman_lengths(freq)
    codes = canonical_codes(lens)
    header_bits = 3 + 5 + 5 + 4 + 19 * 3 + 259 * 4
    dyn_bits = header_bits + sum(f * l for f, l in zip(freq, lens))
    dyn_bytes = (dyn_bits + 3 + 7) // 8 + 4                 # + the closing empty stored block
    stored_bytes = 5 + len(data) + 5
    bw = _Bits(b"\x78\x01" if first else b"")
    if dyn_bytes < stored_bytes:
        bw.put(0, 1)
        bw.put(2, 2)
        bw.put(0, 5)                                        # HLIT: 257 literal/length codes
        bw.put(1, 5)                                        # HDIST: 2 distance codes
        bw.put(15, 4)                                       # HCLEN: all 19 code-length-code lengths
        for s in CL_ORDER:
            bw.put(4 if s < 16 else 0, 3)                   # flat 4-bit code for lengths 0..15, no run-length symbols
        rev4 = lambda v: int(format(v, "04b")[::-1], 2)
        for l in lens:
            bw.put(rev4(l), 4)
        bw.put(rev4(1), 4)
        bw.put(rev4(1), 4)
        # the literals themselves (python ints as the bit accumulator: one big shift per symbol is quadratic, so go in slabs)
        slab = _Bits()
        slabs = []
        for i, v in enumerate(data.tolist()):
            slab.put(codes[v], lens[v])
            if slab.n >= 4096:
                slabs.append(slab)
                slab = _Bits()
        slab.put(codes[256], lens[256])
        slabs.append(slab)
        for sl in slabs:
            bw.put(sl.acc, sl.n)
    else:
        bw.put(0, 8)
        bw.put(len(data), 16)
        bw.put(len(data) ^ 0xFFFF, 16)
        bw.put(int.from_bytes(data.tobytes(), "little"), 8 * len(data))
    bw.put(1 if last else 0, 1)
    bw.put(0, 2)
    bw.align()
    bw.put(0xFFFF0000, 32)
    return bw.bytes()


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))


def encode_png(img: np.ndarray) -> bytes:
    """img uint8 [H, W, C], C in 1..4."""
    H, W, C = img.shape
    row_bytes = 1 + W * C
    if row_bytes > MAX_ROW_BYTES:
        raise ValueError("scanline too long for one chunk")
    R = rows_per_chunk(H, row_bytes)
    out = [b"\x89PNG\r\n\x1a\n", _chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, COLOR_TYPE[C], 0, 0, 0))]
    adler = 1
    n_chunks = (H + R - 1) // R
    for k in range(n_chunks):
        data = _filter_rows(img, k * R, min(H, (k + 1) * R))
        adler = zlib.adler32(data.tobytes(), adler)
        out.append(_chunk(b"IDAT", deflate_chunk(data, k == 0, k == n_chunks - 1)))
    out.append(_chunk(b"IDAT", struct.pack(">I", adler)))
    out.append(_chunk(b"IEND", b""))
    return b"".join(out)


def idat_payload(png: bytes) -> bytes:
    """Concatenated IDAT data (the zlib stream), CRC of every chunk checked."""
    pos, out = 8, []
    while pos < len(png):
        n, tag = struct.unpack(">I4s", png[pos:pos + 8])
        data = png[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", png[pos + 8 + n:pos + 12 + n])
        assert crc == zlib.crc32(tag + data), f"bad CRC in {tag}"
        if tag == b"IDAT":
            out.append(data)
        pos += 12 + n
    return b"".join(out)
