"""The hot segment of the reference's `UnifiedPipeline.__call__` for txt2img
(gyre/pipeline/unified_pipeline.py:1723-1773 signature; :2326-2337 CFG binding; :2341-2350 scheduler choice;
:2432-2483 set_eps_unets / set_timesteps / initial latents / `cscheduler.loop`; :2488-2491 VAE decode and
the image tail), driven through the same objects the reference composes: a guided eps-UNet, a
CommonScheduler and a VAE - all backed by libgyre_b200.

Text encoding is out of scope for this round: the pipeline takes the `[B, 77, C]` text /
uncond embeddings the reference's LPW encoder would produce.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .cfg import B200GuidedUNet
from .common_scheduler import KDiffusionScheduler, SchedulerConfig, build_scheduler
from .modes import EnhancedInpaintMode, EnhancedRunwayInpaintMode, Img2imgMode
from .randtools import batched_randn


@dataclass
class PipelineOutput:
    images: torch.Tensor | None      # [B, 3, H, W] in [0, 1] (output_type "pt"), uint8 NHWC ("uint8"), None ("latent")
    latents: torch.Tensor            # final latents (before the 1/0.18215 scaling)


def generate_latents(generators, batch, in_channels, height, width, sample_size, device, dtype):
    """Txt2imgMode.generateLatents (unified_pipeline.py:193-237): noise is ALWAYS drawn at the UNet's native
    `sample_size` first (one draw per generator), then centre-cropped to, or inserted into the middle of, a
    fresh draw of the requested size - so a seed gives related images across resolutions."""
    h, w = height // 8, width // 8
    shape = (batch, in_channels, h, w)
    mid = batched_randn([batch, in_channels, sample_size, sample_size], generators, device, dtype)
    off2, off3 = (sample_size - h) // 2, (sample_size - w) // 2
    if off2 > 0:
        mid = mid[:, :, off2:off2 + h, :]
    if off3 > 0:
        mid = mid[:, :, :, off3:off3 + w]
    if off2 >= 0 and off3 >= 0:
        return mid.contiguous()
    latents = batched_randn(shape, generators, device, dtype)
    o2, o3 = (latents.shape[2] - mid.shape[2]) // 2, (latents.shape[3] - mid.shape[3]) // 2
    latents[:, :, o2:o2 + mid.shape[2], o3:o3 + mid.shape[3]] = mid
    return latents


class B200Pipeline:
    def __init__(self, unet, vae=None, text_encoder=None):
        self.unet = unet
        self.vae = vae
        self.text_encoder = text_encoder   # B200CLIPTextModel or None: embeddings are passed in
        self.device = unet.device
        self.vae_scale_factor = 8
        self._options = {}
        self.unet_sample_size_override = None   # tests with sub-64 toy UNets

    def get_unet_sample_size(self, unet):
        """unified_pipeline.py:1317-1320: forced minimum of 64."""
        if self.unet_sample_size_override is not None:
            return self.unet_sample_size_override
        return max(64, getattr(unet.config, "sample_size", 64))

    def encode_prompt(self, input_ids, clip_layer="final"):
        """Token ids [B, 77] -> text embeddings [B, 77, C] on the native text encoder, with TextEncoderAltLayer's
        layer choice (text_encoder_alt_layer.py:17-36).  Tokenisation and prompt weighting stay upstream."""
        if self.text_encoder is None:
            raise ValueError("no text encoder attached to this pipeline")
        return self.text_encoder.encode(input_ids, clip_layer)

    def set_options(self, options: dict):
        """Subset of UnifiedPipeline.set_options (unified_pipeline.py:1538-1629) that concerns the hot path."""
        for key, value in options.items():
            if key == "tome":
                # `self.unet.r = int(value)` (:1582-1584); the merge itself is native (gyre_b200.tome_patcher)
                from .tome_patcher import apply_tome
                apply_tome(self.unet)
                self.unet.r = int(value) if not isinstance(value, (tuple, list)) else value
            elif key in ("hires_fix", "grafted_inpaint", "grafted_depth"):
                if value:
                    raise NotImplementedError(f"option {key!r} composes around the boundary and is out of scope")
            else:
                raise ValueError(f"Unknown option {key!r}")
            self._options[key] = value

    @torch.no_grad()
    def __call__(self, prompt_embeds, negative_prompt_embeds, height: int = 512, width: int = 512,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, generator=None,
                 sampler: str = "k_euler_ancestral", scheduler_config: SchedulerConfig | None = None,
                 output_type: str = "pt", callback=None, callback_steps: int = 1, progress_wrapper=None,
                 latents_dtype=torch.float16, return_fp32_latents: bool = False, image=None, mask_image=None,
                 strength: float = 0.8, added_cond_kwargs=None, negative_added_cond_kwargs=None,
                 cfg_execution: str = "parallel") -> PipelineOutput:
        """txt2img (image is None), img2img (image), inpaint (image + mask_image: the 9-channel UNets take the
        EnhancedRunwayInpaintMode path, 4-channel UNets the legacy x0-blend path) - the mode choice of
        unified_pipeline.py:2100-2181.  `image` / `mask_image` are [1, C, H, W] tensors in [0, 1]; the mask is white =
        repaint (the reference's default input convention, preprocess_mask_tensor(inputIs0K1D=True))."""
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if (callback_steps is None) or (not isinstance(callback_steps, int) or callback_steps <= 0):
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps}")
        B = prompt_embeds.shape[0]
        if generator is None:
            raise ValueError("a list of per-sample torch.Generator is required (seeds define the result)")
        generators = list(generator) if isinstance(generator, (list, tuple)) else [generator]
        if B % len(generators) != 0:
            raise ValueError(f"batch {B} is not a multiple of the {len(generators)} generators")
        cfg = self.unet.config
        if image is None and mask_image is not None:
            raise ValueError("Can't pass a mask without an image")
        runway = cfg.in_channels == 9
        if cfg.in_channels not in (4, 9):
            raise NotImplementedError(f"in_channels={cfg.in_channels}: only the 4- and 9-channel UNets are wired up")
        if runway and mask_image is None:
            raise ValueError("an inpainting UNet (in_channels=9) needs image + mask_image")
        if image is not None and self.vae is None:
            raise ValueError("img2img / inpaint need the VAE (encode)")

        if cfg_execution not in ("parallel", "sequential"):
            raise ValueError(f"cfg_execution must be 'parallel' or 'sequential', got {cfg_execution!r}")
        guided = B200GuidedUNet(self.unet, negative_prompt_embeds, prompt_embeds, guidance_scale,
                                parallel=cfg_execution == "parallel")
        if cfg.addition_time_embed_dim:
            if added_cond_kwargs is None:
                raise ValueError("this UNet needs added_cond_kwargs = {text_embeds, time_ids} (text_time conditioning)")
            guided.set_added_cond(negative_added_cond_kwargs or added_cond_kwargs, added_cond_kwargs)
        sched = build_scheduler(sampler, generators, self.device, latents_dtype, callback, callback_steps)
        sched.set_eps_unets([guided])
        sched.use_cuda_graph = bool(getattr(self, "use_cuda_graph", False))
        ts_args = {"strength": min(strength, 1.0)} if image is not None else {}
        sched.set_timesteps(num_inference_steps, prediction_type=cfg.prediction_type,
                            config=scheduler_config or SchedulerConfig(), **ts_args)
        if image is None:
            latents = generate_latents(generators, B, 4, height, width, self.get_unet_sample_size(self.unet),
                                       self.device, latents_dtype)
            latents = sched.prepare_initial_latents(latents)
        else:
            if tuple(image.shape[-2:]) != (height, width):
                raise ValueError(f"image is {tuple(image.shape[-2:])}, expected ({height}, {width})")
            common = dict(pipeline=self, scheduler=sched, generators=generators, image=image,
                          latents_dtype=latents_dtype, batch_total=B)
            if mask_image is None:
                mode = Img2imgMode(strength=strength, **common)
            elif runway:
                mode = EnhancedRunwayInpaintMode(mask_image=mask_image, strength=strength, **common)
            else:
                if not isinstance(sched, KDiffusionScheduler):
                    raise NotImplementedError("legacy (4-channel) inpainting is wired for the k-diffusion samplers")
                mode = EnhancedInpaintMode(mask_image=mask_image, strength=strength, **common)
            latents = mode.generate_latents()
            guided.set_extra_channels(mode.unet_extra_channels())
            blend = mode.x0_blend()
            if blend is not None:
                sched.set_x0_blend(*blend)
        latents = sched.loop(latents, progress_wrapper, out_dtype=torch.float32 if return_fp32_latents else None)
        if output_type == "latent" or self.vae is None:
            return PipelineOutput(images=None, latents=latents)
        z = (1 / self.vae.config.scaling_factor * latents.float()).to(torch.float16)
        img, u8 = self.vae.decode_raw(z.contiguous(), postprocess=True, want_u8=(output_type == "uint8"))
        return PipelineOutput(images=u8 if output_type == "uint8" else img, latents=latents)
