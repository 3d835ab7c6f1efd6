"""Initial-latent construction and UNet-input plumbing of the reference's pipeline modes
(gyre/pipeline/unified_pipeline.py): `Txt2imgMode` (:161-237), `Img2imgMode` (:240-337), the mask helpers of
`MaskProcessorMixin` (:345-397), `EnhancedInpaintMode` (:400-645) and `EnhancedRunwayInpaintMode` (:648-696).

Only what the sampling hot path consumes is here: the latents the loop starts from, the tensors the 9-channel
inpaint UNet is fed every step (mask + masked-image latents) and the x0 blend of the legacy inpaint mode.  The
arithmetic on images runs through the native VAE; mask preparation is a handful of host-side tensor ops executed
once per request (as in the reference).
"""
from __future__ import annotations

import numpy as np
import torch

from .randtools import batched_randn


def downscale_boxop_2d(inp, scale=8, op="max"):
    """unified_pipeline.py:335-343: box min / max over scale x scale pixels."""
    def one(t):
        shape = t.shape[:-1] + (t.shape[-1] // scale, scale)
        return getattr(t.reshape(shape), op)(dim=-1).values
    mid = one(inp)
    return one(mid.transpose(-2, -1)).transpose(-2, -1)


def preprocess_image(tensor):
    """Img2imgMode.preprocess_tensor (:270-281): BCHW, RGB only, [0, 1] -> [-1, 1]."""
    if tensor.ndim == 3:
        tensor = tensor[None, ...]
    tensor = tensor[:, [0, 1, 2]]
    return 2.0 * tensor - 1.0


def preprocess_mask(tensor, input_is_0k1d=True):
    """MaskProcessorMixin.preprocess_mask_tensor (:346-361): 1CHW [0, 1] -> 11HW, 0 = replace, 1 = keep."""
    if tensor.ndim == 3:
        tensor = tensor[None, ...]
    tensor = tensor[:, [0]]
    return 1 - tensor if input_is_0k1d else tensor


def round_mask(mask, threshold=0.5):
    mask = mask.clone()
    mask[mask >= threshold] = 1
    mask[mask < 1] = 0
    return mask


def mask_to_latent_mask(mask):
    """:363-369: 1/8 box-min downsample, 4 channels."""
    return downscale_boxop_2d(mask, 8, "min")[:, [0, 0, 0, 0]]


class Img2imgMode:
    """Encode -> per-generator posterior sample -> x 0.18215 -> add the start timestep's noise (:283-332)."""

    def __init__(self, pipeline, scheduler, generators, image, latents_dtype, batch_total, strength, max_strength=1.0):
        if strength < 0 or strength > max_strength:
            raise ValueError(f"The value of strength should in [0.0, {max_strength:.1f}] but is {strength}")
        self.pipeline, self.scheduler, self.generators = pipeline, scheduler, list(generators)
        self.device = pipeline.device
        self.latents_dtype = latents_dtype
        self.batch_total = batch_total
        self.image = preprocess_image(image)

    def convert_to_latents(self, image, mask=None):
        image = image.to(device=self.device, dtype=torch.float16)
        if mask is not None:
            image = image * (mask.to(self.device) > 0.5)
        dist = self.pipeline.vae.encode(image).latent_dist
        latents = torch.cat([dist.sample(generator=g) for g in self.generators], dim=0)
        return 0.18215 * latents.to(self.device, self.latents_dtype)

    def build_initial_latents(self):
        return self.convert_to_latents(self.image)

    def add_initial_noise(self, latents):
        self.image_noise = batched_randn(latents.shape, self.generators, self.device, self.latents_dtype)
        return self.scheduler.add_noise(latents, self.image_noise, self.scheduler.start_timestep).to(latents.dtype)

    def generate_latents(self):
        return self.add_initial_noise(self.build_initial_latents())

    # what the mode contributes to the per-step path
    def unet_extra_channels(self):
        return None

    def x0_blend(self):
        return None


class EnhancedInpaintMode(Img2imgMode):
    """Legacy inpainting with a 4-channel UNet: the original latents are blended back into the x0 prediction
    while `latent_blend_mask > u` (:400-645)."""

    def __init__(self, mask_image, strength, **kw):
        if strength < 0 or strength > 2:
            raise ValueError(f"The value of strength should in [0.0, 2.0] but is {strength}")
        self.fill_with_shaped_noise = strength >= 1.0
        self.shaped_noise_strength = min(2 - strength, 1)
        super().__init__(strength=min(strength, 1), **kw)
        self.mask = preprocess_mask(mask_image).to(device=self.device, dtype=self.latents_dtype)
        high_mask = round_mask(self.mask, 0.001)
        self.init_latents_orig = self.convert_to_latents(self.image, high_mask)
        self.latent_mask = torch.cat([mask_to_latent_mask(self.mask)] * self.batch_total)
        self.latent_high_mask = round_mask(self.latent_mask, 0.001)
        self.latent_low_mask = round_mask(self.latent_mask, 0.999)
        self.latent_blend_mask = self.latent_mask * 1

    def fill_with_shaped_noise_mode5(self, init_latents):
        """`_fillWithShapedNoise(noise_mode=5)` (:462-592): the replaceable area is filled with pixels drawn (numpy
        RNG seeded from the sample's generator) from the kept area of the same channel, mixed with plain noise."""
        masked = init_latents * self.latent_high_mask
        batch_noise = []
        for generator, split in zip(self.generators, masked.split(1)):
            npseed = torch.randint(low=0, high=torch.iinfo(torch.int32).max, size=[1], generator=generator,
                                   device=generator.device, dtype=torch.int32).cpu()
            npgen = np.random.default_rng(npseed.numpy())
            keep = self.latent_high_mask[[0], [0]].ge(0.5)
            channels = []
            for channel in split.split(1, dim=1):
                good = channel.masked_select(keep)
                mixed = npgen.choice(good.float().cpu().numpy(), tuple(channel.shape))
                channels.append(torch.from_numpy(mixed).to(split.device).to(split.dtype))
            noise = torch.zeros_like(split).to(generator.device).normal_(generator=generator).to(split.device)
            noise = noise * (1 - self.shaped_noise_strength) + torch.cat(channels, dim=1) * self.shaped_noise_strength
            batch_noise.append(noise)
        noise = torch.cat(batch_noise, dim=0)
        return init_latents * self.latent_mask + noise * (1 - self.latent_mask)

    def generate_latents(self):
        init = self.build_initial_latents()
        if self.fill_with_shaped_noise:
            init = self.fill_with_shaped_noise_mode5(init)
        return self.add_initial_noise(init)

    def x0_blend(self):
        return self.init_latents_orig.float().contiguous(), self.latent_blend_mask.float().contiguous()


class EnhancedRunwayInpaintMode(EnhancedInpaintMode):
    """Inpainting UNets (in_channels = 9): every step's UNet input is cat([latents, mask (1 = repaint),
    masked-image latents]) and no x0 blending happens (:648-696)."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self.inpaint_mask = 1 - self.latent_high_mask[:, [0]]
        self.masked_image_latents = self.init_latents_orig

    def unet_extra_channels(self):
        return torch.cat([self.inpaint_mask, self.masked_image_latents], dim=1).to(torch.float16).contiguous()

    def x0_blend(self):
        return None
