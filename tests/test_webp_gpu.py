"""gyre_b200_webp_encode against its CPU restatement (oracle/webp.py, pinned by libwebp decoding it) BYTE FOR BYTE, and against
the decoder directly at the headline batch."""
import io

import numpy as np
import pytest
import torch

from oracle import webp as owebp
from oracle.safety import synthetic_image

pytestmark = pytest.mark.gpu


def _encode(imgs_u8):
    from gyre_b200.images import to_webp_bytes
    return to_webp_bytes(torch.from_numpy(np.ascontiguousarray(imgs_u8)).cuda())


def _decode(data, C):
    from PIL import Image
    im = Image.open(io.BytesIO(data))
    im.load()
    return np.asarray(im.convert("RGBA" if C == 4 else "RGB"))


CASES = [("synthetic 64x96", lambda r: synthetic_image(64, 96)),
         ("noise 33x17", lambda r: r.integers(0, 256, (33, 17, 3), dtype=np.uint8)),
         ("constant 40x40", lambda r: np.zeros((40, 40, 3), np.uint8)),
         ("1x1", lambda r: np.full((1, 1, 3), 7, np.uint8)),
         ("rgba 20x30", lambda r: r.integers(0, 256, (20, 30, 4), dtype=np.uint8)),
         ("two values 50x30", lambda r: (r.integers(0, 2, (50, 30, 1), dtype=np.uint8) * 200).repeat(3, axis=2)),
         ("one row 1x700", lambda r: synthetic_image(1, 700)),
         ("one column 600x1", lambda r: synthetic_image(600, 1)),
         ("several chunks 70x300", lambda r: synthetic_image(70, 300)),
         ("skewed (length limit) 128x128", lambda r: (np.minimum(r.geometric(0.6, (128, 128, 3)) - 1, 255).cumsum(axis=1) % 256).astype(np.uint8))]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_webp_bytes_equal_oracle(name, make):
    img = make(np.random.default_rng(3)).astype(np.uint8)
    got = _encode(img[None])[0]
    ref = owebp.encode_webp(img)
    assert len(got) == len(ref), (len(got), len(ref))
    assert got == ref, f"first difference at byte {next(i for i, (a, b) in enumerate(zip(got, ref)) if a != b)}"
    assert np.array_equal(_decode(got, img.shape[2]), img)


def test_webp_batch_full_size_roundtrip():
    rng = np.random.default_rng(7)
    y, x = np.mgrid[0:512, 0:512]
    imgs = np.stack([np.stack([127 + 100 * np.sin(x / (20 + 5 * i) + c) * np.cos(y / 31 - c) + rng.normal(0, 3 + i, (512, 512))
                               for c in range(3)], -1).clip(0, 255).astype(np.uint8) for i in range(8)])
    imgs[7] = synthetic_image(512, 512)
    files = _encode(imgs)
    assert len(files) == 8
    for i, f in enumerate(files):
        assert np.array_equal(_decode(f, 3), imgs[i]), i
    assert files[2] == owebp.encode_webp(imgs[2])
    assert _encode(imgs[3:4])[0] == files[3]                      # independent of the batch neighbours


def test_to_webp_bytes_mirrors_reference_signature():
    from gyre_b200.images import add_text_chunk_to_webp_bytes, to_webp_bytes
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 3, 24, 40, generator=g)
    files = to_webp_bytes(x.cuda())
    u8 = (x * 255).round().to(torch.uint8).permute(0, 2, 3, 1).numpy()
    for f, ref in zip(files, u8):
        assert np.array_equal(_decode(f, 3), ref)
    assert to_webp_bytes(x[0].cuda())[0] == files[0]
    rgba = torch.rand(1, 4, 8, 8, generator=g)
    assert np.array_equal(_decode(to_webp_bytes(rgba.cuda())[0], 4), (rgba * 255).round().to(torch.uint8)[0].permute(1, 2, 0).numpy())
    grey = torch.rand(1, 1, 8, 8, generator=g)
    assert np.array_equal(_decode(to_webp_bytes(grey.cuda())[0], 3)[..., 0], (grey * 255).round().to(torch.uint8)[0, 0].numpy())
    tagged = add_text_chunk_to_webp_bytes(files[0], b"ICMT", "steps=50")
    assert np.array_equal(_decode(tagged, 3), u8[0])


def test_pipeline_webp_output_decodes_to_uint8_output():
    from oracle.unet import UNetConfig, synth_params, unet_param_shapes
    from oracle.vae import VAEConfig, vae_param_shapes
    from gyre_b200.pipeline import B200Pipeline
    from gyre_b200.unet import B200UNet
    from gyre_b200.vae import B200VAE
    ucfg, vcfg = UNetConfig.tiny(), VAEConfig.tiny()
    pipe = B200Pipeline(B200UNet(ucfg).load_state_dict(synth_params(unet_param_shapes(ucfg), seed=1234)),
                        B200VAE(vcfg).load_state_dict(synth_params(vae_param_shapes(vcfg), seed=4321)))
    pipe.unet_sample_size_override = 16
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 77, ucfg.cross_attention_dim, generator=g).half().cuda()
    unc = torch.randn(2, 77, ucfg.cross_attention_dim, generator=g).half().cuda()
    kw = dict(height=128, width=128, num_inference_steps=3, sampler="k_euler")
    u8 = pipe(emb, unc, generator=[torch.Generator("cpu").manual_seed(s) for s in (5, 6)], output_type="uint8", **kw).images
    webp = pipe(emb, unc, generator=[torch.Generator("cpu").manual_seed(s) for s in (5, 6)], output_type="webp", **kw).images
    assert isinstance(webp, list) and len(webp) == 2 and all(isinstance(f, bytes) and f[8:12] == b"WEBP" for f in webp)
    for f, ref in zip(webp, u8.cpu().numpy()):
        assert np.array_equal(_decode(f, 3), ref)
