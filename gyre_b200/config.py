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
        """Accepts this class, a dict, or any object with the same attribute names (e.g. the oracle's or a
        diffusers config)."""
        if isinstance(cfg, cls):
            return cfg
        get = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
        out = cls()
        for f in cls.__dataclass_fields__:
            v = get(f)
            if v is not None:
                setattr(out, f, tuple(v) if isinstance(v, (list, tuple)) else v)
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
