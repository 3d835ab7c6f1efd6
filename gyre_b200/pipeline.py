"""The hot segment of the reference's `UnifiedPipeline.__call__` for txt2img
(gyre/pipeline/unified_pipeline.py:1723-1773 signature; :2326-2337 CFG binding; :2341-2350 scheduler choice;
:2432-2483 set_eps_unets / set_timesteps / initial latents / `cscheduler.loop`; :2488-2491 VAE decode and
the image tail), driven through the same objects the reference composes: a guided eps-UNet, a
CommonScheduler and a VAE - all backed by libgyre_b200.

The mode tree (ModeTreeRoot / Node / Leaf, unified_pipeline.py:1065-1217) is kept: a request is one leaf, or a
GraftUnets / HiresUnetWrapper composition of leaves, each leaf with its own guided UNet, mode and k-unet.

The pipeline takes the `[B, 77, C]` text / uncond embeddings (the LPW encoder of SURVEY.md 8f1 produces them, see
gyre_b200.text_encoder).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from .cfg import B200GuidedUNet
from .common_scheduler import SAMPLERS, KDiffusionScheduler, SchedulerConfig, build_scheduler
from .modes import EnhancedInpaintMode, EnhancedRunwayInpaintMode, Img2imgMode
from .randtools import batched_randn


@dataclass
class PipelineOutput:
    images: object                   # [B, 3, H, W] in [0, 1] (output_type "pt"), uint8 NHWC ("uint8"), list of PNG / lossless
                                     # WebP files as bytes ("png" / "webp": images.toPngBytes / toWebpBytes of the reference,
                                     # encoded on the device), None ("latent")
    latents: torch.Tensor            # final latents (before the 1/0.18215 scaling)
    nsfw_content_detected: list | None = None   # per image, from the safety checker (unified_pipeline.py:2514-2524)


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


class _Txt2imgLeafMode:
    """Txt2imgMode (unified_pipeline.py:161-237) as the object a mode-tree leaf holds."""

    def __init__(self, pipeline, scheduler, generators, unet, height, width, latents_dtype, batch_total):
        self.args = (generators, batch_total, 4, height, width, pipeline.get_unet_sample_size(unet), pipeline.device,
                     latents_dtype)
        self.scheduler = scheduler

    def generate_latents(self):
        return self.scheduler.prepare_initial_latents(generate_latents(*self.args))

    def unet_extra_channels(self):
        return None

    def x0_blend(self):
        return None


class _Leaf:
    """ModeTreeLeaf (unified_pipeline.py:1183-1217): the options of one UNet evaluation path."""

    def __init__(self, **opts):
        self.opts = opts
        self.mode = None
        self.guided = None
        self.k_unet = None

    def clone(self, **overrides):
        return _Leaf(**{**self.opts, **overrides})

    @property
    def leaves(self):
        return [self]

    def collapse(self):
        return self.k_unet

    def initial_latents(self):
        return self.mode.generate_latents()

    def split_result(self, result):
        return result


class _Node:
    """ModeTreeNode (unified_pipeline.py:1140-1180): two sub-trees merged by HiresUnetWrapper / GraftUnets."""

    def __init__(self, left, right, merger, **kwargs):
        self.left, self.right, self.merger, self.kwargs = left, right, merger, kwargs

    def clone(self, **overrides):
        return _Node(self.left.clone(**overrides), self.right.clone(**overrides), self.merger, **self.kwargs)

    @property
    def leaves(self):
        return self.left.leaves + self.right.leaves

    def collapse(self):
        return self.merger(self.left.collapse(), self.right.collapse(), **self.kwargs)

    def initial_latents(self):
        return self.merger.merge_initial_latents(self.left.initial_latents(), self.right.initial_latents())

    def split_result(self, result):
        return self.merger.split_result(self.left.split_result(result), self.right.split_result(result))


class B200Pipeline:
    def __init__(self, unet, vae=None, text_encoder=None, inpaint_unet=None, tokenizer=None, depth_unet=None,
                 safety_checker=None, feature_extractor=None):
        self.unet = unet
        # B200SafetyChecker + B200FeatureExtractor, or None: no check, every image reported clean
        self.safety_checker = safety_checker
        self.feature_extractor = feature_extractor
        if safety_checker is not None and feature_extractor is None:
            from .safety_checker import B200FeatureExtractor
            self.feature_extractor = B200FeatureExtractor(size=safety_checker.config.image_size, device=unet.device)
        self.depth_unet = depth_unet       # 5-channel depth2img UNet (unified_pipeline.py:1334, 1974-2013), or None
        self.tokenizer = tokenizer         # the caller's CLIPTokenizer (or any object with its call surface) for `prompt=`
        self.inpaint_unet = inpaint_unet   # 9-channel UNet of the same family (unified_pipeline.py:2059-2062), or None
        self.vae = vae
        self.text_encoder = text_encoder   # B200CLIPTextModel or None: embeddings are passed in
        self.device = unet.device
        self.vae_scale_factor = 8
        self._options = {}
        self.unet_sample_size_override = None   # tests with sub-64 toy UNets
        # engine defaults of the reference (unified_pipeline.py:1362-1373)
        self._grafted_inpaint = False
        self._grafted_depth = False
        self._hires_fix = True
        self._hires_threshold_fraction = 0.0333
        self._hires_oos_fraction = 0.6
        self._hires_image_oos_fraction = 1.0
        self._text_embedding_layer = "final"

    def get_unet_sample_size(self, unet):
        """unified_pipeline.py:1317-1320: forced minimum of 64."""
        if self.unet_sample_size_override is not None:
            return self.unet_sample_size_override
        return max(64, getattr(unet.config, "sample_size", 64))

    def get_unet_pixel_size(self, unet):
        return self.get_unet_sample_size(unet) * self.vae_scale_factor

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
            elif key == "hires_fix":
                self._hires_fix = bool(value)
            elif key == "hires":
                for subkey, subval in value.items():
                    if subkey == "enable":
                        self._hires_fix = bool(subval)
                    elif subkey == "threshold_fraction":
                        self._hires_threshold_fraction = float(subval)
                    elif subkey == "oos_fraction":
                        self._hires_oos_fraction = float(subval)
                    elif subkey == "image_oos_fraction":
                        self._hires_image_oos_fraction = float(subval)
                    else:
                        raise ValueError(f"Unknown option {subkey}: {subval} passed as part of hires settings")
            elif key == "grafted_inpaint":
                self._grafted_inpaint = value if isinstance(value, dict) else bool(value)
            elif key == "grafted_depth":
                self._grafted_depth = value if isinstance(value, dict) else bool(value)
            elif key == "text_embedding_layer":
                # the CLIP layer prompts are embedded at unless the request says otherwise (:1624-1625, 2228-2243)
                self._text_embedding_layer = value
            elif key == "xformers":
                pass        # attention always runs on the native flash kernel: nothing to switch (:1567-1579)
            elif key == "graft_factor":
                print("Graft Factor is no longer used. Please remove it from your engines.yaml.")       # (:1562-1566)
            elif key == "structured_diffusion":
                if value:
                    print("structured diffusion is deprecated")                                       # (:1589-1590)
            elif key in ("clip", "clip_vae_grad"):
                # CLIP guidance differentiates a CLIP loss through the UNet and the VAE (unet/clipguided.py:301-338):
                # inference-only kernels cannot serve it
                raise NotImplementedError("CLIP guidance options: not available on the native path (needs autograd through the UNet / VAE)")
            else:
                raise ValueError(f"Unknown option {key}: {value} passed to UnifiedPipeline")
            self._options[key] = value

    def embed_prompts(self, prompt, negative_prompt=None, max_embeddings_multiples: int = 3, clip_layer="final"):
        """`LPWTextEmbedding(...).get_embeddings(prompt, uncond_prompt)` (unified_pipeline.py:2269-2304): weighted text and
        uncond embeddings [B, 77 * k, C] from prompt strings (or pre-parsed (text, weight) lists) on the native CLIP."""
        if self.text_encoder is None or self.tokenizer is None:
            raise ValueError("text prompts need a text encoder and a tokenizer attached to the pipeline")
        from .lpw_text_embedding import LPWTextEmbedding
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        if negative_prompt is None:
            negative_prompt = [""] * len(prompts)
        negs = [negative_prompt] * len(prompts) if isinstance(negative_prompt, str) else list(negative_prompt)
        if len(negs) != len(prompts):
            raise ValueError(f"`negative_prompt` has batch size {len(negs)}, `prompt` has {len(prompts)}")
        calc = LPWTextEmbedding(max_embeddings_multiples, tokenizer=self.tokenizer, text_encoder=self.text_encoder,
                                uncond_encoder=self.text_encoder, device=self.device, clip_layer=clip_layer)
        return calc.get_embeddings(prompts, negs)

    @torch.no_grad()
    def __call__(self, prompt_embeds=None, negative_prompt_embeds=None, height: int = 512, width: int = 512,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, generator=None,
                 sampler: str = "k_euler_ancestral", scheduler_config: SchedulerConfig | None = None,
                 output_type: str = "pt", callback=None, callback_steps: int = 1, progress_wrapper=None,
                 latents_dtype=torch.float16, return_fp32_latents: bool = False, image=None, mask_image=None,
                 strength: float = 0.8, added_cond_kwargs=None, negative_added_cond_kwargs=None,
                 cfg_execution: str = "parallel", hires_fix: bool | None = None,
                 hires_oos_fraction: float | None = None, outmask_image=None, prompt=None, negative_prompt=None,
                 max_embeddings_multiples: int = 3, clip_layer=None, depth_map=None,
                 run_safety_checker: bool = True, hints=None, depth_image=None) -> PipelineOutput:
        """txt2img (image is None), img2img (image), inpaint (image + mask_image: the 9-channel UNets take the
        EnhancedRunwayInpaintMode path, 4-channel UNets the legacy x0-blend path) - the mode choice of
        unified_pipeline.py:2055-2066 - optionally grafted (inpaint UNet early, main UNet late, :2069-2098) and, for
        requests above the UNet's native size, doubled by the hires fix (:2100-2181).  `image` / `mask_image` are
        [1, C, H, W] tensors in [0, 1]; the mask is white = repaint (the reference's default input convention,
        preprocess_mask_tensor(inputIs0K1D=True))."""
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if prompt_embeds is None:
            if prompt is None:
                raise ValueError("pass `prompt` (text, needs the text encoder + tokenizer) or `prompt_embeds`")
            prompt_embeds, negative_prompt_embeds = self.embed_prompts(
                prompt, negative_prompt, max_embeddings_multiples, self._text_embedding_layer if clip_layer is None else clip_layer)
        elif prompt is not None:
            raise ValueError("pass either `prompt` or `prompt_embeds`, not both")
        if (callback_steps is None) or (not isinstance(callback_steps, int) or callback_steps <= 0):
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps}")
        B = prompt_embeds.shape[0]
        if generator is None:
            raise ValueError("a list of per-sample torch.Generator is required (seeds define the result)")
        generators = list(generator) if isinstance(generator, (list, tuple)) else [generator]
        if B % len(generators) != 0:
            raise ValueError(f"batch {B} is not a multiple of the {len(generators)} generators")
        if image is None and mask_image is not None:
            raise ValueError("Can't pass a mask without an image")
        if image is not None and self.vae is None:
            raise ValueError("img2img / inpaint need the VAE (encode)")
        if image is not None and tuple(image.shape[-2:]) != (height, width):
            raise ValueError(f"image is {tuple(image.shape[-2:])}, expected ({height}, {width})")
        if cfg_execution not in ("parallel", "sequential"):
            raise ValueError(f"cfg_execution must be 'parallel' or 'sequential', got {cfg_execution!r}")
        if hires_fix is None:
            hires_fix = self._hires_fix
        if hires_oos_fraction is None:
            hires_oos_fraction = self._hires_image_oos_fraction if image is not None else self._hires_oos_fraction

        # ---- the main mode leaf (unified_pipeline.py:2055-2066)
        main_unet = self.unet
        if mask_image is not None and self.inpaint_unet is not None:
            main_unet = self.inpaint_unet
        # a depth hint goes to the depth UNet when the engine has one and no mask is given (:1974-2013); `depth_map` is the
        # hint already normalised to [-1, 1] at latent resolution ([1 | B, 1, H / 8, W / 8]) - estimating and resizing it is
        # the hinter's job upstream
        if depth_image is not None:
            # a depth hint at image resolution (unified_pipeline.py:2004-2013): first channel, scaled to the latent grid with
            # images.resize(.., 1 / 8, sharpness=2) - lanczos3 without antialiasing - and mapped to [-1, 1]
            if depth_map is not None:
                raise ValueError("pass either depth_map (latent resolution, [-1, 1]) or depth_image ([0, 1]), not both")
            from .images import resize
            d = depth_image if depth_image.ndim == 4 else depth_image[None]
            depth_map = 2.0 * resize(d[:, [0]].to(self.device), (1 / 8, 1 / 8), sharpness=2) - 1.0
        if depth_map is not None:
            if self.depth_unet is None or mask_image is not None:
                raise EnvironmentError("a depth map needs a depth UNet and cannot be combined with a mask")
            if tuple(depth_map.shape[-2:]) != (height // 8, width // 8) or depth_map.shape[1] != 1:
                raise ValueError(f"depth_map is {tuple(depth_map.shape)}, expected [1 | B, 1, {height // 8}, {width // 8}]")
            main_unet = self.depth_unet
        cfg = main_unet.config
        if cfg.in_channels not in (4, 5, 9):
            raise NotImplementedError(f"in_channels={cfg.in_channels}: only the 4-, 5- and 9-channel UNets are wired up")
        if cfg.in_channels == 9 and mask_image is None:
            raise ValueError("an inpainting UNet (in_channels=9) needs image + mask_image")
        if (cfg.in_channels == 5) != (depth_map is not None):
            raise ValueError("a depth UNet (in_channels=5) needs depth_map, and only it takes one")
        if mask_image is not None:
            kind = "runway" if cfg.in_channels == 9 else "inpaint"
        else:
            kind = "img2img" if image is not None else "txt2img"
        # `hints`: B200ControlnetHint / B200T2iHint objects (gyre_b200.hints; UnifiedPipelineHint.for_model upstream,
        # unified_pipeline.py:2018-2034): ControlNets run at every UNet call, adapter states once per request
        tree = _Leaf(kind=kind, unet=main_unet, height=height, width=width, image=image, mask_image=mask_image,
                     depth_map=depth_map, hints=list(hints or []))

        # ---- graft: inpaint UNet for the early steps, the main UNet with the legacy x0 blend for the late ones (:2069-2098)
        if kind == "runway" and self._grafted_inpaint and main_unet is self.inpaint_unet and self.unet is not main_unet:
            from .graft import GraftUnets
            blend = self._grafted_inpaint if isinstance(self._grafted_inpaint, dict) else {}
            tree = _Node(tree.clone(), tree.clone(kind="inpaint", unet=self.unet), GraftUnets, generators=generators,
                         blend=blend, rand_dtype=latents_dtype)

        # ---- grafted depth: the depth UNet for the early steps, the main UNet (no depth input) for the late ones
        if depth_map is not None and self._grafted_depth and self.unet is not main_unet:
            from .graft import GraftUnets
            blend = self._grafted_depth if isinstance(self._grafted_depth, dict) else {}
            tree = _Node(tree.clone(), tree.clone(unet=self.unet, depth_map=None), GraftUnets, generators=generators,
                         blend=blend, rand_dtype=latents_dtype)

        # ---- hires fix: a natural-size twin of every leaf, cross-blended with the full-size one (:2100-2181)
        if hires_fix:
            unet_pixel_size = self.get_unet_pixel_size(self.unet)
            sample_size = self.get_unet_sample_size(self.unet)
            threshold = math.floor(unet_pixel_size * (1 + self._hires_threshold_fraction))
            too_small = width < unet_pixel_size or height < unet_pixel_size
            if not too_small and not (width <= threshold and height <= threshold):
                if sampler not in SAMPLERS or SAMPLERS[sampler][0] is not KDiffusionScheduler:
                    raise ValueError("Can't use Diffuser schedulers with Hires fix. "
                                     "Either use a K-Diffusion scheduler or disable Hires fix.")
                from .hires_fix import HiresUnetWrapper

                def to_natural(t):
                    return None if t is None else HiresUnetWrapper.image_to_natural(
                        unet_pixel_size, t.to(self.device), oos_fraction=hires_oos_fraction)

                def depth_to_natural(leaf_opts):
                    d = leaf_opts.get("depth_map")
                    return None if d is None else HiresUnetWrapper.image_to_natural(sample_size, d.to(self.device),
                                                                                    oos_fraction=hires_oos_fraction)
                natural = tree.clone(width=unet_pixel_size, height=unet_pixel_size, image=to_natural(image),
                                     mask_image=to_natural(mask_image))
                for leaf in natural.leaves:          # per leaf: a grafted twin has no depth map
                    leaf.opts["depth_map"] = depth_to_natural(leaf.opts)
                    # the natural-size twin sees the hint image scaled like the init image (:2148-2158)
                    leaf.opts["hints"] = [type(h)(h.model, to_natural(h.image.float()), None, h.weight, h.soft_injection, h.cfg_only)
                                          for h in leaf.opts.get("hints") or []]
                tree = _Node(natural, tree, HiresUnetWrapper, generators=generators,
                             natural_size=[sample_size, sample_size], oos_fraction=hires_oos_fraction,
                             latent_debugger=None, rand_dtype=latents_dtype)
        leaves = tree.leaves

        # ---- one guided eps-UNet per leaf (CFG + embeddings, :2326-2337); leaves of one UNet share its K/V context
        first_of = {}
        for leaf in leaves:
            u = leaf.opts["unet"]
            g = B200GuidedUNet(u, negative_prompt_embeds, prompt_embeds, guidance_scale, parallel=cfg_execution == "parallel")
            if u.config.addition_time_embed_dim:
                if added_cond_kwargs is None:
                    raise ValueError("this UNet needs added_cond_kwargs = {text_embeds, time_ids} (text_time conditioning)")
                g.set_added_cond(negative_added_cond_kwargs or added_cond_kwargs, added_cond_kwargs)
            g.ctx_owner = first_of.setdefault(id(u), g)
            g.set_hints(leaf.opts.get("hints"))
            leaf.guided = g
        sched = build_scheduler(sampler, generators, self.device, latents_dtype, callback, callback_steps)
        sched.set_eps_unets([leaf.guided for leaf in leaves])
        # (hints allocate per call: the whole-loop graph is for the plain path)
        sched.use_cuda_graph = bool(getattr(self, "use_cuda_graph", False)) and not hints
        ts_args = {"strength": min(strength, 1.0)} if image is not None else {}
        sched.set_timesteps(num_inference_steps, prediction_type=cfg.prediction_type,
                            config=scheduler_config or SchedulerConfig(), **ts_args)

        # ---- modes, in leaf order (construction already consumes generator draws: posterior samples of the encodes)
        for leaf in leaves:
            o = leaf.opts
            common = dict(pipeline=self, scheduler=sched, generators=generators, latents_dtype=latents_dtype, batch_total=B)
            if o["kind"] == "txt2img":
                leaf.mode = _Txt2imgLeafMode(unet=o["unet"], height=o["height"], width=o["width"], **common)
            elif o["kind"] == "img2img":
                leaf.mode = Img2imgMode(image=o["image"], strength=strength, **common)
            elif o["kind"] == "runway":
                leaf.mode = EnhancedRunwayInpaintMode(image=o["image"], mask_image=o["mask_image"], strength=strength, **common)
            else:
                if not isinstance(sched, KDiffusionScheduler):
                    raise NotImplementedError("legacy (4-channel) inpainting is wired for the k-diffusion samplers")
                leaf.mode = EnhancedInpaintMode(image=o["image"], mask_image=o["mask_image"], strength=strength, **common)
            extra = leaf.mode.unet_extra_channels()
            if o.get("depth_map") is not None:
                # UnetWithExtraChannels(unet, depth_map) (unified_pipeline.py:2306-2310, unet/core.py:21-37): un-scaled
                extra = o["depth_map"].to(self.device, torch.float16).expand(B, -1, -1, -1).contiguous()
            leaf.guided.set_extra_channels(extra)
        if len(leaves) == 1:
            blend = leaves[0].mode.x0_blend()
            if blend is not None:
                sched.set_x0_blend(*blend)
        else:
            # `leaf.k_unet = leaf.mode.wrap_k_unet(cscheduler.unets[i])`, then `cscheduler.unet = mode_tree.collapse()`
            # (:2461-2471)
            for i, leaf in enumerate(leaves):
                leaf.k_unet = sched.unets[i]
                blend = leaf.mode.x0_blend()
                if blend is not None:
                    leaf.k_unet.blend = (blend[0].float().contiguous(), blend[1].float().contiguous())
            sched.unet = tree.collapse()
        latents = tree.initial_latents()
        latents = sched.loop(latents, progress_wrapper, out_dtype=torch.float32 if return_fp32_latents else None)
        latents = tree.split_result(latents)
        if output_type == "latent" or self.vae is None:
            return PipelineOutput(images=None, latents=latents)
        z = (1 / self.vae.config.scaling_factor * latents.float()).to(torch.float16)
        outpaint = image is not None and outmask_image is not None
        img, u8 = self.vae.decode_raw(z.contiguous(), postprocess=True, want_u8=(output_type in ("uint8", "png", "webp") and not outpaint))
        if outpaint:
            # unified_pipeline.py:2493-2510: histogram-match the result to the source around it, mix the source back in
            from .images import match_histograms_outpaint, to_uint8_nhwc
            img = match_histograms_outpaint(img, image, outmask_image)
            u8 = to_uint8_nhwc(img) if output_type in ("uint8", "png", "webp") else None
        if run_safety_checker and self.safety_checker is not None:
            # unified_pipeline.py:2514-2522: the checker looks at the 8-bit image (numpy_to_pil's quantisation) through the
            # CLIP feature extractor; here both stay on the device and only the 20 scores per image come back
            clip_input = self.feature_extractor(u8 if u8 is not None else img, return_tensors="pt").pixel_values
            _, has_nsfw = self.safety_checker(images=img, clip_input=clip_input.to(latents_dtype))
        else:
            has_nsfw = [False] * img.shape[0]
        if output_type == "png":
            from .images import to_png_bytes
            return PipelineOutput(images=to_png_bytes(u8), latents=latents, nsfw_content_detected=has_nsfw)
        if output_type == "webp":      # what services/generate.py:73-76 picks when the client accepts image/webp
            from .images import to_webp_bytes
            return PipelineOutput(images=to_webp_bytes(u8), latents=latents, nsfw_content_detected=has_nsfw)
        return PipelineOutput(images=u8 if output_type == "uint8" else img, latents=latents, nsfw_content_detected=has_nsfw)
