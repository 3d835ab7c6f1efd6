"""The lossless-WebP encoder's CPU restatement (oracle/webp.py) pinned by a decoder: Pillow's libwebp opens what it emits and
returns exactly the input pixels (RGB and RGBA, 1x1 to several chunks, constant / two-valued / noisy images)."""
import io

import numpy as np
import pytest

from oracle import webp as owebp
from oracle.safety import synthetic_image

CASES = [("synthetic 64x96", lambda r: synthetic_image(64, 96)),
         ("noise 33x17", lambda r: r.integers(0, 256, (33, 17, 3), dtype=np.uint8)),
         ("constant 40x40", lambda r: np.zeros((40, 40, 3), np.uint8)),
         ("1x1", lambda r: np.full((1, 1, 3), 7, np.uint8)),
         ("rgba 20x30", lambda r: r.integers(0, 256, (20, 30, 4), dtype=np.uint8)),
         ("two values 50x30", lambda r: (r.integers(0, 2, (50, 30, 1), dtype=np.uint8) * 200).repeat(3, axis=2)),
         ("one row 1x700", lambda r: synthetic_image(1, 700)),
         ("one column 600x1", lambda r: synthetic_image(600, 1)),
         ("several chunks 70x300", lambda r: synthetic_image(70, 300))]


def _decode(data, C):
    Image = pytest.importorskip("PIL.Image")
    from PIL import features
    if not features.check("webp"):
        pytest.skip("Pillow was built without WebP")
    im = Image.open(io.BytesIO(data))
    im.load()
    return np.asarray(im.convert("RGBA" if C == 4 else "RGB"))


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_oracle_webp_decodes_with_libwebp(name, make):
    img = make(np.random.default_rng(3)).astype(np.uint8)
    data = owebp.encode_webp(img)
    assert data[:4] == b"RIFF" and data[8:16] == b"WEBPVP8L" and len(data) % 2 == 0
    assert int.from_bytes(data[4:8], "little") == len(data) - 8
    assert np.array_equal(_decode(data, img.shape[2]), img)


def test_text_chunk_insertion_keeps_the_image():
    from gyre_b200.images import add_text_chunk_to_webp_bytes
    img = synthetic_image(16, 24)
    data = add_text_chunk_to_webp_bytes(owebp.encode_webp(img), b"ICMT", "steps=50")
    assert data[12:16] == b"VP8X" and b"ICMT" in data and int.from_bytes(data[4:8], "little") == len(data) - 8
    assert np.array_equal(_decode(data, 3), img)
