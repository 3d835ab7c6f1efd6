"""gyre_b200_png_encode against its CPU restatement (oracle/png.py, itself pinned by the decoders in test_png_cpu.py) BYTE
FOR BYTE, and against the decoders directly at sizes the oracle is not run at."""
import io
import zlib

import numpy as np
import pytest
import torch

from oracle import png as opng
from oracle.safety import synthetic_image

pytestmark = pytest.mark.gpu


def _encode(imgs_u8):
    from gyre_b200.images import to_png_bytes
    return to_png_bytes(torch.from_numpy(np.ascontiguousarray(imgs_u8)).cuda())


def _decode_pil(data, shape):
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data))).reshape(shape)


CASES = [("synthetic 64x96 rgb", lambda r: synthetic_image(64, 96)),
         ("noise 33x17 rgb", lambda r: r.integers(0, 256, (33, 17, 3), dtype=np.uint8)),
         ("constant 40x40 rgb", lambda r: np.zeros((40, 40, 3), np.uint8)),
         ("1x1 grey", lambda r: np.full((1, 1, 1), 7, np.uint8)),
         ("rgba 20x30", lambda r: r.integers(0, 256, (20, 30, 4), dtype=np.uint8)),
         ("grey 50x30 smooth", lambda r: (np.add.outer(np.arange(50), np.arange(30)) % 256).astype(np.uint8)[..., None]),
         ("grey+alpha 9x11", lambda r: r.integers(0, 256, (9, 11, 2), dtype=np.uint8)),
         ("several chunks 70x300 rgb", lambda r: synthetic_image(70, 300)),
         ("one row 1x500 rgb", lambda r: synthetic_image(1, 500)),
         ("one column 300x1 rgb", lambda r: synthetic_image(300, 1)),
         ("skewed histogram 64x64 grey", lambda r: ((np.minimum(r.geometric(0.5, (64, 64, 1)) - 1, 255) * 7) % 256).astype(np.uint8)),
         ("wide 3x10000 rgb", lambda r: synthetic_image(3, 10000))]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_png_bytes_equal_oracle(name, make):
    img = make(np.random.default_rng(3)).astype(np.uint8)
    got = _encode(img[None])[0]
    ref = opng.encode_png(img)
    assert len(got) == len(ref), (len(got), len(ref))
    assert got == ref, f"first difference at byte {next(i for i, (a, b) in enumerate(zip(got, ref)) if a != b)}"
    assert np.array_equal(_decode_pil(got, img.shape), img)


def test_png_batch_full_size_roundtrip():
    """8 x 512x512 RGB (the headline batch): every file decodes to its image with Pillow and torchvision, image 0 equals the
    oracle's bytes, and files are independent of their batch neighbours."""
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(7)
    y, x = np.mgrid[0:512, 0:512]
    imgs = np.stack([np.stack([127 + 100 * np.sin(x / (20 + 5 * i) + c) * np.cos(y / 31 - c) + rng.normal(0, 3 + i, (512, 512))
                               for c in range(3)], -1).clip(0, 255).astype(np.uint8) for i in range(8)])
    imgs[7] = synthetic_image(512, 512)
    files = _encode(imgs)
    assert len(files) == 8
    for i, f in enumerate(files):
        raw = zlib.decompress(opng.idat_payload(f))
        assert len(raw) == 512 * (1 + 512 * 3)
        assert np.array_equal(_decode_pil(f, imgs[i].shape), imgs[i])
        dec = tv.io.decode_image(torch.frombuffer(bytearray(f), dtype=torch.uint8), tv.io.image.ImageReadMode.RGB)
        assert np.array_equal(dec.permute(1, 2, 0).numpy(), imgs[i])
    assert files[0] == opng.encode_png(imgs[0])
    assert _encode(imgs[3:4])[0] == files[3]
    ref_size = tv.io.encode_png(torch.from_numpy(imgs[0]).permute(2, 0, 1).contiguous()).numel()
    assert len(files[0]) < 1.05 * ref_size, (len(files[0]), ref_size)


def test_to_png_bytes_mirrors_reference_signature():
    """toPngBytes(tensor): float [B, C, H, W] / [C, H, W] in [0, 1], quantised (x * 255).round(); 1, 3 or 4 channels."""
    from gyre_b200.images import add_text_chunk_to_png_bytes, to_png_bytes
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 3, 24, 40, generator=g)
    files = to_png_bytes(x.cuda())
    u8 = (x * 255).round().to(torch.uint8).permute(0, 2, 3, 1).numpy()
    for f, ref in zip(files, u8):
        assert np.array_equal(_decode_pil(f, ref.shape), ref)
    one = to_png_bytes(x[0].cuda())                      # [C, H, W] like the reference's ndim == 3 branch
    assert len(one) == 1 and one[0] == files[0]
    rgba = torch.rand(1, 4, 8, 8, generator=g)
    dec = _decode_pil(to_png_bytes(rgba.cuda())[0], (8, 8, 4))
    assert np.array_equal(dec, (rgba * 255).round().to(torch.uint8)[0].permute(1, 2, 0).numpy())
    assert to_png_bytes(torch.rand(1, 5, 8, 8).cuda()) == []
    from PIL import Image
    im = Image.open(io.BytesIO(add_text_chunk_to_png_bytes(files[0], "generation_parameters", "steps=50")))
    im.load()
    assert im.text["generation_parameters"] == "steps=50"


def test_png_refuses_overlong_scanlines():
    from gyre_b200 import _native as N
    from gyre_b200.images import encode_png_u8
    with pytest.raises(N.NativeError):
        encode_png_u8(torch.zeros(1, 2, 11000, 3, dtype=torch.uint8).cuda())
    with pytest.raises(ValueError):
        encode_png_u8(torch.zeros(1, 3, 8, 8).cuda())


def test_pipeline_png_output_decodes_to_uint8_output():
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
    png = pipe(emb, unc, generator=[torch.Generator("cpu").manual_seed(s) for s in (5, 6)], output_type="png", **kw).images
    assert isinstance(png, list) and len(png) == 2 and all(isinstance(p, bytes) for p in png)
    for f, ref in zip(png, u8.cpu().numpy()):
        assert np.array_equal(_decode_pil(f, ref.shape), ref)
