"""The PNG encoder's CPU restatement (oracle/png.py) pinned by the decoders: zlib inflates the concatenated IDATs to the
filtered scanlines, every chunk CRC and the Adler-32 check out, and Pillow and the reference's own decoder
(`torchvision.io.decode_image`, gyre/images.py:85-90 fromPngBytes) return exactly the pixels that the reference's encoder
(`torchvision.io.encode_png`, gyre/images.py:93-111 toPngBytes) round-trips."""
import io
import zlib

import numpy as np
import pytest
import torch

from oracle import png as opng
from oracle.safety import synthetic_image

CASES = [("synthetic 64x96 rgb", lambda r: synthetic_image(64, 96)),
         ("noise 33x17 rgb (stored blocks)", lambda r: r.integers(0, 256, (33, 17, 3), dtype=np.uint8)),
         ("constant 40x40 rgb", lambda r: np.zeros((40, 40, 3), np.uint8)),
         ("1x1 grey", lambda r: np.full((1, 1, 1), 7, np.uint8)),
         ("rgba 20x30", lambda r: r.integers(0, 256, (20, 30, 4), dtype=np.uint8)),
         ("grey 50x30 smooth", lambda r: (np.add.outer(np.arange(50), np.arange(30)) % 256).astype(np.uint8)[..., None]),
         ("grey+alpha 9x11", lambda r: r.integers(0, 256, (9, 11, 2), dtype=np.uint8)),
         ("several chunks 70x300 rgb", lambda r: synthetic_image(70, 300)),
         ("one row 1x500 rgb", lambda r: synthetic_image(1, 500)),
         ("one column 300x1 rgb", lambda r: synthetic_image(300, 1))]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_oracle_png_decodes_everywhere(name, make):
    Image = pytest.importorskip("PIL.Image")
    img = make(np.random.default_rng(3))
    H, W, C = img.shape
    data = opng.encode_png(img)
    raw = zlib.decompress(opng.idat_payload(data))                 # CRCs asserted inside, Adler-32 by zlib
    assert len(raw) == H * (1 + W * C)
    assert all(raw[r * (1 + W * C)] <= 4 for r in range(H))       # filter types
    pil = np.asarray(Image.open(io.BytesIO(data)))
    assert np.array_equal(pil.reshape(H, W, C), img)
    tv = pytest.importorskip("torchvision")
    dec = tv.io.decode_image(torch.frombuffer(bytearray(data), dtype=torch.uint8), tv.io.image.ImageReadMode.UNCHANGED)
    assert np.array_equal(dec.permute(1, 2, 0).numpy(), img)
    if C in (1, 3):
        ref_file = tv.io.encode_png(torch.from_numpy(img).permute(2, 0, 1).contiguous())       # the reference's encoder
        ref_dec = tv.io.decode_image(ref_file, tv.io.image.ImageReadMode.UNCHANGED)
        assert torch.equal(dec, ref_dec)


def test_oracle_png_size_against_libpng():
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(0)
    y, x = np.mgrid[0:256, 0:256]
    img = np.stack([127 + 100 * np.sin(x / 40 + c) * np.cos(y / 31 - c) + rng.normal(0, 4, (256, 256)) for c in range(3)],
                   -1).clip(0, 255).astype(np.uint8)
    mine = len(opng.encode_png(img))
    ref = tv.io.encode_png(torch.from_numpy(img).permute(2, 0, 1).contiguous()).numel()
    assert mine < 1.05 * ref, (mine, ref)


def test_huffman_lengths_are_complete_and_limited():
    rng = np.random.default_rng(1)
    fib = [1, 1]
    while len(fib) < 40:
        fib.append(fib[-1] + fib[-2])
    for freq in ([1, 1], [5, 0, 0, 1], fib[:30], fib, rng.integers(0, 50, 257).tolist() + [0], [1] * 257, [10 ** 6] + [1] * 256):
        lens = opng.huffman_lengths(freq)
        used = [l for l in lens if l]
        assert len(used) == sum(1 for f in freq if f) and max(used) <= 15
        assert sum(2 ** (15 - l) for l in used) == 2 ** 15, freq[:8]      # complete code: zlib rejects anything else
        # more frequent symbols never get longer codes
        order = sorted((f, s) for s, f in enumerate(freq) if f)
        assert all(lens[order[i][1]] >= lens[order[i + 1][1]] for i in range(len(order) - 1))
        codes = opng.canonical_codes(lens)
        strs = {format(codes[s], f"0{l}b")[::-1] for s, l in enumerate(lens) if l}
        assert len(strs) == len(used) and not any(a != b and b.startswith(a) for a in strs for b in strs)


def test_text_chunk_insertion_matches_reference_expression():
    from gyre_b200.images import add_text_chunk_to_png_bytes
    Image = pytest.importorskip("PIL.Image")
    data = opng.encode_png(synthetic_image(8, 8))
    out = add_text_chunk_to_png_bytes(data, "generation_parameters", "seed=1")
    opng.idat_payload(out)                                         # all CRCs, the new chunk's included
    im = Image.open(io.BytesIO(out))
    im.load()
    assert im.text["generation_parameters"] == "seed=1"
    assert out.index(b"tEXt") < out.rindex(b"IEND") and out.endswith(data[-12:])
