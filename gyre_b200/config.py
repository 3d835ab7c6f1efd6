"""Architecture hyper-parameters of the models on the hot path (reference:
gyre/ldm_config/v1-inference.yaml:29-67, v2-inference-v.yaml; diffusers-0.16 config names)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: tuple = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    # diffusers-0.16: `attention_head_dim` is the NUMBER of heads per level
    num_heads: tuple = (8, 8, 8, 8)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    attn_levels: tuple = (True, True, True, False)
    sample_size: int = 64
    prediction_type: str = "epsilon"
    upcast_attention: bool = False
    # SDXL-style topologies (BASELINE config 5): BasicTransformerBlocks per Transformer2DModel per level, and the
    # text_time additional conditioning (pooled text embedding ++ 256-dim sinusoids of the 6 time ids)
    transformer_layers_per_block: tuple = (1, 1, 1, 1)
    addition_time_embed_dim: int = 0
    projection_class_embeddings_input_dim: int = 0

    @staticmethod
    def sdxl(**kw):
        return UNetConfig(block_out_channels=(320, 640, 1280), num_heads=(5, 10, 20), attn_levels=(False, True, True),
                          transformer_layers_per_block=(1, 2, 10), cross_attention_dim=2048, use_linear_projection=True,
                          sample_size=128, addition_time_embed_dim=256, projection_class_embeddings_input_dim=2816, **kw)

    @staticmethod
    def sd15(**kw):
        return UNetConfig(**kw)

    @staticmethod
    def sd15_inpaint(**kw):
        return UNetConfig(in_channels=9, **kw)

    @staticmethod
    def sd21_v(**kw):
        return UNetConfig(num_heads=(5, 10, 20, 20), cross_attention_dim=1024, use_linear_projection=True,
                          sample_size=96, prediction_type="v_prediction", upcast_attention=True, **kw)

    @staticmethod
    def tiny(**kw):
        d = dict(block_out_channels=(64, 128, 256, 256), num_heads=(4, 4, 4, 4), cross_attention_dim=64,
                 sample_size=16)
        d.update(kw)
        return UNetConfig(**d)

    @classmethod
    def from_any(cls, cfg):
        """Accepts this class, a dict / object with this class's field names (the oracle's config), or a diffusers
        `UNet2DConditionModel.config` (a dict-like with the diffusers names).  The diffusers names that differ from
        ours are mapped explicitly (reference construction sites: gyre/pipeline/ckpt_utils.py:280-304,
        gyre/pipeline/controlnet/models.py:100-131):

        * `attention_head_dim` (int or per-level list) is the NUMBER of heads in diffusers 0.16 -> `num_heads`
          (`num_attention_heads`, when a newer config carries it, wins);
        * `down_block_types` -> `attn_levels` ("CrossAttnDownBlock2D" levels have transformers);
        * `only_cross_attention`, `dual_cross_attention`, `class_embed_type`, `num_class_embeds`, a non-default
          `act_fn` / `time_embedding_type` / `resnet_time_scale_shift` are not built: they raise.
        A config from which the head count or the attention levels cannot be resolved raises instead of silently
        running every attention layer with the wrong head split."""
        if isinstance(cfg, cls):
            return cfg
        if isinstance(cfg, dict):
            has = lambda k: k in cfg and cfg[k] is not None          # noqa: E731
            get = lambda k, d=None: cfg[k] if has(k) else d          # noqa: E731
        else:
            has = lambda k: getattr(cfg, k, None) is not None        # noqa: E731
            get = lambda k, d=None: getattr(cfg, k, d) if has(k) else d   # noqa: E731
        out = cls()
        for f in cls.__dataclass_fields__:
            v = get(f)
            if v is not None:
                setattr(out, f, tuple(v) if isinstance(v, (list, tuple)) else v)
        n_levels = len(out.block_out_channels)

        def per_level(v, name):
            if isinstance(v, (list, tuple)):
                if len(v) != n_levels:
                    raise ValueError(f"UNet config: {name} has {len(v)} entries for {n_levels} levels")
                return tuple(int(x) for x in v)
            return (int(v),) * n_levels

        diffusers_style = has("attention_head_dim") or has("down_block_types") or has("num_attention_heads")
        if has("num_attention_heads"):
            out.num_heads = per_level(get("num_attention_heads"), "num_attention_heads")
        elif has("attention_head_dim") and not has("num_heads"):
            out.num_heads = per_level(get("attention_head_dim"), "attention_head_dim")
        elif diffusers_style and not has("num_heads"):
            raise ValueError("UNet config: cannot resolve the number of attention heads "
                             "(no attention_head_dim / num_attention_heads / num_heads)")
        if has("down_block_types") and not has("attn_levels"):
            types = list(get("down_block_types"))
            if len(types) != n_levels:
                raise ValueError(f"UNet config: {len(types)} down_block_types for {n_levels} levels")
            known = {"CrossAttnDownBlock2D": True, "DownBlock2D": False}
            for t in types:
                if t not in known:
                    raise NotImplementedError(f"UNet config: down block type {t!r} is not built")
            out.attn_levels = tuple(known[t] for t in types)
            if has("up_block_types"):
                ups = list(get("up_block_types"))
                want = ["CrossAttnUpBlock2D" if a else "UpBlock2D" for a in reversed(out.attn_levels)]
                if ups != want:
                    raise NotImplementedError(f"UNet config: up_block_types {ups} do not mirror the down blocks")
        elif diffusers_style and not has("attn_levels"):
            raise ValueError("UNet config: cannot resolve which levels carry attention (no down_block_types / attn_levels)")
        if has("transformer_layers_per_block"):
            out.transformer_layers_per_block = per_level(get("transformer_layers_per_block"), "transformer_layers_per_block")
        if len(out.num_heads) != n_levels or len(out.attn_levels) != n_levels:
            raise ValueError("UNet config: num_heads / attn_levels must have one entry per level")
        for c, h_, a_ in zip(out.block_out_channels, out.num_heads, out.attn_levels):
            if a_ and (h_ <= 0 or c % h_ != 0):
                raise ValueError(f"UNet config: {c} channels do not split into {h_} heads")
        oca = get("only_cross_attention", False)
        if (any(oca) if isinstance(oca, (list, tuple)) else bool(oca)) or get("dual_cross_attention", False):
            raise NotImplementedError("UNet config: only_cross_attention / dual_cross_attention are not built")
        if get("class_embed_type") is not None or get("num_class_embeds") is not None:
            raise NotImplementedError("UNet config: class embeddings are not built")
        if get("act_fn", "silu") not in ("silu", "swish"):
            raise NotImplementedError(f"UNet config: act_fn {get('act_fn')!r} is not built")
        if get("resnet_time_scale_shift", "default") != "default" or get("time_embedding_type", "positional") != "positional":
            raise NotImplementedError("UNet config: only default resnet_time_scale_shift / positional time embedding are built")
        aet = get("addition_embed_type")
        if aet not in (None, "text_time"):
            raise NotImplementedError(f"UNet config: addition_embed_type {aet!r} is not built")
        return out


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: tuple = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215

    @staticmethod
    def sd():
        return VAEConfig()

    @staticmethod
    def tiny():
        return VAEConfig(block_out_channels=(64, 64, 128, 128))

    @classmethod
    def from_any(cls, cfg):
        if isinstance(cfg, cls):
            return cfg
        get = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
        out = cls()
        for f in cls.__dataclass_fields__:
            v = get(f)
            if v is not None:
                setattr(out, f, tuple(v) if isinstance(v, (list, tuple)) else v)
        return out
