"""Import the reference's *vendored* k-diffusion and ToMe sources from /root/reference (this container
only; the GPU box has no /root/reference).  Same stub trick gyre uses (gyre/src/__init__.py:52-71)."""
import importlib.util
import os
import sys
import types

REF = os.environ.get("GYRE_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "gyre/src/k-diffusion/k_diffusion"))


def _load(pkg, base, name):
    if pkg not in sys.modules:
        m = types.ModuleType(pkg)
        m.__path__ = [base]
        sys.modules[pkg] = m
    full = f"{pkg}.{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(base, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    setattr(sys.modules[pkg], name, mod)
    return mod


def k_diffusion():
    for stub in ("torchsde", "torchdiffeq"):
        if stub not in sys.modules:
            sys.modules[stub] = types.ModuleType(stub)
    sys.modules["torchdiffeq"].odeint = None
    base = os.path.join(REF, "gyre/src/k-diffusion/k_diffusion")
    utils = _load("k_diffusion", base, "utils")
    sampling = _load("k_diffusion", base, "sampling")
    external = _load("k_diffusion", base, "external")
    return utils, sampling, external


def tome_merge():
    base = os.path.join(REF, "nonfree/ToMe/tome")
    return _load("tome", base, "merge")


def gyre_dpmpp_2m():
    p = os.path.join(REF, "gyre/pipeline/schedulers/sample_dpmpp_2m.py")
    if "_gyre_sample_dpmpp_2m" in sys.modules:
        return sys.modules["_gyre_sample_dpmpp_2m"]
    spec = importlib.util.spec_from_file_location("_gyre_sample_dpmpp_2m", p)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_gyre_sample_dpmpp_2m"] = mod
    spec.loader.exec_module(mod)
    return mod


def gyre_ddim():
    """The in-tree DDIM copy (gyre/pipeline/schedulers/scheduling_ddim.py) with a minimal `diffusers` stub
    (ConfigMixin / register_to_config / BaseOutput / SchedulerMixin are plumbing, not arithmetic)."""
    import functools
    import inspect

    if "diffusers" not in sys.modules:
        d = types.ModuleType("diffusers")
        cu = types.ModuleType("diffusers.configuration_utils")
        ut = types.ModuleType("diffusers.utils")
        sc = types.ModuleType("diffusers.schedulers")
        su = types.ModuleType("diffusers.schedulers.scheduling_utils")

        class _Cfg(dict):
            __getattr__ = dict.__getitem__

        class ConfigMixin:
            pass

        def register_to_config(init):
            @functools.wraps(init)
            def inner(self, *a, **kw):
                sig = inspect.signature(init)
                b = sig.bind(self, *a, **kw)
                b.apply_defaults()
                self.config = _Cfg({k: v for k, v in b.arguments.items() if k != "self"})
                init(self, *a, **kw)
            return inner

        class BaseOutput:
            pass

        class SchedulerMixin:
            pass

        cu.ConfigMixin, cu.register_to_config = ConfigMixin, register_to_config
        ut.BaseOutput, ut.deprecate = BaseOutput, (lambda *a, **k: None)
        su.SchedulerMixin = SchedulerMixin
        for n, m in (("diffusers", d), ("diffusers.configuration_utils", cu), ("diffusers.utils", ut),
                     ("diffusers.schedulers", sc), ("diffusers.schedulers.scheduling_utils", su)):
            sys.modules[n] = m
    p = os.path.join(REF, "gyre/pipeline/schedulers/scheduling_ddim.py")
    if "_gyre_scheduling_ddim" in sys.modules:
        return sys.modules["_gyre_scheduling_ddim"]
    spec = importlib.util.spec_from_file_location("_gyre_scheduling_ddim", p)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_gyre_scheduling_ddim"] = mod          # (inspect.getmodule, used by gyre/patching.py, goes through sys.modules)
    spec.loader.exec_module(mod)
    return mod


def gyre_pipeline_pure():
    """The pure-torch pieces of gyre/pipeline that the hot path is wrapped in, loaded file by file under synthetic
    `gyre`, `gyre.pipeline`, `gyre.pipeline.unet` packages (the real package __init__ files pull in diffusers, which is
    absent): randtools.py, unet/types.py, unet/cfg.py, unet/core.py.  Returns (randtools, types, cfg, core)."""
    for name, rel in (("gyre", "gyre"), ("gyre.pipeline", "gyre/pipeline"), ("gyre.pipeline.unet", "gyre/pipeline/unet")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF, rel)]
            sys.modules[name] = m
    rt = _load("gyre.pipeline", os.path.join(REF, "gyre/pipeline"), "randtools")
    base = os.path.join(REF, "gyre/pipeline/unet")
    ty = _load("gyre.pipeline.unet", base, "types")
    cfg = _load("gyre.pipeline.unet", base, "cfg")
    core = _load("gyre.pipeline.unet", base, "core")
    return rt, ty, cfg, core


def gyre_hires():
    """gyre/pipeline/unet/hires_fix.py, unet/graft.py and easing.py with the REAL vendored ResizeRight
    (gyre/src/ResizeRight, through gyre/resize_right.py).  The absent third-party `easing_functions` package
    (easing-functions ~= 1.0.4) is stood in for by the oracle's restated curves (oracle/hires.py), so `Easing.interp`
    itself is the reference's.  Returns (hires_fix, graft, easing)."""
    from oracle import hires as ohires
    gyre_pipeline_pure()
    if "easing_functions" not in sys.modules:
        ef = types.ModuleType("easing_functions")
        efe = types.ModuleType("easing_functions.easing")
        for n in ("EasingBase", "LinearInOut", "QuadEaseInOut", "CubicEaseInOut", "QuarticEaseInOut", "QuinticEaseInOut",
                  "SineEaseInOut", "CircularEaseInOut", "ExponentialEaseInOut"):
            setattr(efe, n, getattr(ohires, n))
        ef.easing = efe
        sys.modules["easing_functions"], sys.modules["easing_functions.easing"] = ef, efe
    for name, rel in (("gyre.src", "gyre/src"), ("gyre.src.ResizeRight", "gyre/src/ResizeRight")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF, rel)]
            sys.modules[name] = m
    rr_base = os.path.join(REF, "gyre/src/ResizeRight")
    im = _load("gyre.src.ResizeRight", rr_base, "interp_methods")
    sys.modules["interp_methods"] = im
    _load("gyre.src.ResizeRight", rr_base, "resize_right")
    _load("gyre", os.path.join(REF, "gyre"), "resize_right")
    easing = _load("gyre.pipeline", os.path.join(REF, "gyre/pipeline"), "easing")
    base = os.path.join(REF, "gyre/pipeline/unet")
    hf = _load("gyre.pipeline.unet", base, "hires_fix")
    gr = _load("gyre.pipeline.unet", base, "graft")
    return hf, gr, easing


def gyre_lpw():
    """gyre/pipeline/text_embedding/lpw_text_embedding.py (imports the INSTALLED transformers CLIP classes for type
    annotations only) under a synthetic `gyre.pipeline.text_embedding` package."""
    gyre_pipeline_pure()
    name = "gyre.pipeline.text_embedding"
    base = os.path.join(REF, "gyre/pipeline/text_embedding")
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = [base]
        sys.modules[name] = m
    _load(name, base, "text_embedding")
    return _load(name, base, "lpw_text_embedding")


def gyre_t2i_adapter():
    """gyre/pipeline/t2i_adapter/adapter.py (pure torch; its package __init__ and models.py import diffusers, so the file
    and its utils.py are loaded under a synthetic package)."""
    gyre_pipeline_pure()
    name = "gyre.pipeline.t2i_adapter"
    base = os.path.join(REF, "gyre/pipeline/t2i_adapter")
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = [base]
        sys.modules[name] = m
    _load(name, base, "utils")
    return _load(name, base, "adapter")


def gyre_safety_checkers():
    """gyre/pipeline/safety_checkers.py: imports only torch / numpy / transformers (installed), loaded as a lone file."""
    p = os.path.join(REF, "gyre/pipeline/safety_checkers.py")
    name = "_gyre_safety_checkers"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, p)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def gyre_controlnet():
    """gyre/pipeline/controlnet/models.py (the in-tree ControlNetModel: ControlNetConditioningEmbedding, zero convolutions,
    construction and forward wiring) with the absent `diffusers` building blocks stood in for by nn.Modules that hold their
    parameters under the diffusers names and evaluate the oracle's restated blocks (oracle/unet.py: resnet_block,
    transformer_2d, timestep_embedding).  What this makes checkable is everything the FILE states - not the blocks."""
    import torch
    from torch import nn
    from oracle import unet as ou
    gyre_ddim()                                    # installs the diffusers.configuration_utils / utils stubs
    d = sys.modules["diffusers"]
    ut = sys.modules["diffusers.utils"]
    if not hasattr(ut, "logging"):
        lg = types.ModuleType("diffusers.utils.logging")
        import logging as _pylog
        lg.get_logger = _pylog.getLogger
        ut.logging = lg
        sys.modules["diffusers.utils.logging"] = lg
    if "diffusers.models.unet_2d_blocks" not in sys.modules:
        models = types.ModuleType("diffusers.models")
        ca = types.ModuleType("diffusers.models.cross_attention")
        emb = types.ModuleType("diffusers.models.embeddings")
        mu = types.ModuleType("diffusers.models.modeling_utils")
        blk = types.ModuleType("diffusers.models.unet_2d_blocks")

        class AttnProcessor:
            pass

        class ModelMixin(nn.Module):
            @property
            def dtype(self):
                return next(self.parameters()).dtype

        class Timesteps(nn.Module):
            def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
                super().__init__()
                assert flip_sin_to_cos and downscale_freq_shift == 0          # what the oracle restates (SD configs)
                self.num_channels = num_channels

            def forward(self, t):
                return ou.timestep_embedding(t, self.num_channels)

        class TimestepEmbedding(nn.Module):
            def __init__(self, in_channels, time_embed_dim, act_fn="silu"):
                super().__init__()
                assert act_fn == "silu"
                self.linear_1 = nn.Linear(in_channels, time_embed_dim)
                self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

            def forward(self, sample, condition=None):
                assert condition is None
                return self.linear_2(torch.nn.functional.silu(self.linear_1(sample)))

        def _register(mod, shapes, prefix):
            """Parameters of an oracle block as (nested) module attributes so that state_dict() has the diffusers names."""
            for k, shp in shapes.items():
                assert k.startswith(prefix + ".")
                parts = k[len(prefix) + 1:].split(".")
                cur = mod
                for p_ in parts[:-1]:
                    if not hasattr(cur, p_):
                        setattr(cur, p_, nn.Module())
                    cur = getattr(cur, p_)
                setattr(cur, parts[-1], nn.Parameter(torch.randn(shp) / max(1, int(torch.tensor(shp[1:]).prod())) ** 0.5))

        class _Block(nn.Module):
            def params(self, prefix):
                return {f"{prefix}.{k}": v for k, v in self.state_dict().items()}

        class DownBlock2D(_Block):
            has_cross_attention = False

            def __init__(self, num_layers, in_channels, out_channels, temb_channels, add_downsample, resnet_eps, resnet_groups,
                         cross_attention_dim=None, heads=None, linear=False, **_):
                super().__init__()
                self.n, self.eps, self.groups, self.heads, self.linear = num_layers, resnet_eps, resnet_groups, heads, linear
                self.add_downsample = add_downsample
                shapes = {}
                cin = in_channels
                for j in range(num_layers):
                    shapes.update(ou._resnet_keys(f"b.resnets.{j}", cin, out_channels, temb_channels))
                    cin = out_channels
                    if self.has_cross_attention:
                        shapes.update(ou._transformer_keys(f"b.attentions.{j}", out_channels, cross_attention_dim, linear, 1))
                if add_downsample:
                    shapes["b.downsamplers.0.conv.weight"] = (out_channels, out_channels, 3, 3)
                    shapes["b.downsamplers.0.conv.bias"] = (out_channels,)
                _register(self, shapes, "b")

            def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                        cross_attention_kwargs=None):
                assert attention_mask is None and cross_attention_kwargs is None
                P = self.params("b")
                out = ()
                h = hidden_states
                for j in range(self.n):
                    h = ou.resnet_block(P, f"b.resnets.{j}", h, temb, self.groups, self.eps)
                    if self.has_cross_attention:
                        h = ou.transformer_2d(P, f"b.attentions.{j}", h, encoder_hidden_states, self.heads, self.groups,
                                              self.linear, 0)
                    out += (h,)
                if self.add_downsample:
                    h = torch.nn.functional.conv2d(h, P["b.downsamplers.0.conv.weight"], P["b.downsamplers.0.conv.bias"],
                                                   stride=2, padding=1)
                    out += (h,)
                return h, out

        class CrossAttnDownBlock2D(DownBlock2D):
            has_cross_attention = True

        def get_down_block(down_block_type, num_layers, in_channels, out_channels, temb_channels, add_downsample, resnet_eps,
                           resnet_act_fn, resnet_groups, cross_attention_dim, attn_num_head_channels, downsample_padding,
                           use_linear_projection, only_cross_attention, upcast_attention, resnet_time_scale_shift):
            assert resnet_act_fn == "silu" and downsample_padding == 1 and not only_cross_attention and not upcast_attention
            assert resnet_time_scale_shift == "default"
            cls = {"CrossAttnDownBlock2D": CrossAttnDownBlock2D, "DownBlock2D": DownBlock2D}[down_block_type]
            return cls(num_layers, in_channels, out_channels, temb_channels, add_downsample, resnet_eps, resnet_groups,
                       cross_attention_dim=cross_attention_dim, heads=attn_num_head_channels, linear=use_linear_projection)

        class UNetMidBlock2DCrossAttn(_Block):
            def __init__(self, in_channels, temb_channels, resnet_eps, resnet_act_fn, output_scale_factor,
                         resnet_time_scale_shift, cross_attention_dim, attn_num_head_channels, resnet_groups,
                         use_linear_projection, upcast_attention):
                super().__init__()
                assert output_scale_factor == 1 and resnet_act_fn == "silu" and not upcast_attention
                self.eps, self.groups, self.heads, self.linear = resnet_eps, resnet_groups, attn_num_head_channels, use_linear_projection
                shapes = {}
                shapes.update(ou._resnet_keys("b.resnets.0", in_channels, in_channels, temb_channels))
                shapes.update(ou._transformer_keys("b.attentions.0", in_channels, cross_attention_dim, use_linear_projection, 1))
                shapes.update(ou._resnet_keys("b.resnets.1", in_channels, in_channels, temb_channels))
                _register(self, shapes, "b")

            def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                        cross_attention_kwargs=None):
                assert attention_mask is None and cross_attention_kwargs is None
                P = self.params("b")
                h = ou.resnet_block(P, "b.resnets.0", hidden_states, temb, self.groups, self.eps)
                h = ou.transformer_2d(P, "b.attentions.0", h, encoder_hidden_states, self.heads, self.groups, self.linear, 0)
                return ou.resnet_block(P, "b.resnets.1", h, temb, self.groups, self.eps)

        ca.AttnProcessor = AttnProcessor
        emb.Timesteps, emb.TimestepEmbedding = Timesteps, TimestepEmbedding
        mu.ModelMixin = ModelMixin
        blk.CrossAttnDownBlock2D, blk.DownBlock2D = CrossAttnDownBlock2D, DownBlock2D
        blk.UNetMidBlock2DCrossAttn, blk.get_down_block = UNetMidBlock2DCrossAttn, get_down_block
        d.models = models
        for n, m in (("diffusers.models", models), ("diffusers.models.cross_attention", ca), ("diffusers.models.embeddings", emb),
                     ("diffusers.models.modeling_utils", mu), ("diffusers.models.unet_2d_blocks", blk)):
            sys.modules[n] = m
    p = os.path.join(REF, "gyre/pipeline/controlnet/models.py")
    name = "_gyre_controlnet_models"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, p)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def gyre_attention_modules():
    """gyre/pipeline/models/memory_efficient_cross_attention.py (a pure nn.Module) and
    nonfree/tome_memory_efficient_cross_attention.py (its ToMe variant, a subclass of diffusers' CrossAttention) with
      * `xformers.ops.memory_efficient_attention(q, k, v, attn_bias=None, op=...)` stood in for by its published definition,
        softmax(q k^T / sqrt(d)) v over [B * heads, N, d] (xformers is absent);
      * `diffusers.models.attention.CrossAttention` stood in for by an nn.Module with the attributes the subclass reads
        (to_q / to_k / to_v without bias, to_out = [Linear, Dropout], heads, dim_head);
      * the REAL vendored ToMe merge (nonfree/ToMe/tome/merge.py) and tome/utils.parse_r.
    Returns (memory_efficient_cross_attention module, tome_memory_efficient_cross_attention module)."""
    import torch
    from torch import nn
    if "xformers" not in sys.modules or not hasattr(sys.modules["xformers"], "ops"):
        xf = types.ModuleType("xformers")
        ops = types.ModuleType("xformers.ops")

        def memory_efficient_attention(q, k, v, attn_bias=None, op=None):
            assert attn_bias is None
            s = (q @ k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
            return torch.softmax(s, dim=-1) @ v
        ops.memory_efficient_attention = memory_efficient_attention
        xf.ops = ops
        sys.modules["xformers"], sys.modules["xformers.ops"] = xf, ops
    gyre_ddim()
    if "diffusers.models" not in sys.modules:
        sys.modules["diffusers.models"] = types.ModuleType("diffusers.models")
        sys.modules["diffusers"].models = sys.modules["diffusers.models"]
    if "diffusers.models.attention" not in sys.modules:
        att = types.ModuleType("diffusers.models.attention")

        class CrossAttention(nn.Module):
            def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0):
                super().__init__()
                inner = dim_head * heads
                cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
                self.heads, self.dim_head = heads, dim_head
                self.to_q = nn.Linear(query_dim, inner, bias=False)
                self.to_k = nn.Linear(cross_attention_dim, inner, bias=False)
                self.to_v = nn.Linear(cross_attention_dim, inner, bias=False)
                # the subclass calls `self.to_out(out)`: the CrossAttention it was written against held an nn.Sequential
                self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))
        att.CrossAttention = CrossAttention
        sys.modules["diffusers.models.attention"] = att
        sys.modules["diffusers.models"].attention = att
    tome_merge()
    _load("tome", os.path.join(REF, "nonfree/ToMe/tome"), "utils")
    out = []
    for name, rel in (("_gyre_mem_eff_attn", "gyre/pipeline/models/memory_efficient_cross_attention.py"),
                      ("_gyre_tome_mem_eff_attn", "nonfree/tome_memory_efficient_cross_attention.py")):
        if name not in sys.modules:
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
        out.append(sys.modules[name])
    return tuple(out)


# ------------------------------------------------------------------------------------------------------------------
# gyre/pipeline/unified_pipeline.py itself.  The file imports half of gyre and a dozen absent third-party packages at module
# level, but the classes the hot path is wrapped in (Txt2imgMode / Img2imgMode / EnhancedInpaintMode /
# EnhancedRunwayInpaintMode, UnifiedPipelineHint_*, the mode tree) only use torch once they are defined.  So: real gyre files
# from /root/reference under synthetic package objects (no package __init__ side effects), and a LAST-RESORT meta-path finder
# that turns every module that cannot be found into a permissive stand-in (CamelCase attributes become empty classes, the rest
# MagicMocks).  Nothing in the stand-ins computes anything: whatever arithmetic runs afterwards is the reference's own.
class _AutoStubBase:
    def __init__(self, *a, **k):
        pass

    def __init_subclass__(cls, **kwargs):
        pass

    def __class_getitem__(cls, item):
        return cls

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        from unittest.mock import MagicMock
        return MagicMock(name=name)


def _auto_attr(modname, name):
    from unittest.mock import MagicMock
    if name.startswith("__"):
        raise AttributeError(name)
    if name[:1].isupper() and not name.isupper():
        return type(name, (_AutoStubBase,), {"__module__": modname})
    return MagicMock(name=f"{modname}.{name}")


class _AutoStubModule(types.ModuleType):
    def __getattr__(self, name):
        v = _auto_attr(self.__name__, name)
        setattr(self, name, v)
        return v


class _AutoStubFinder:
    """Appended to sys.meta_path: consulted only after every real finder has failed."""

    # top-level packages the reference imports that this image does not have (or, for `gyre`, generated / optional files)
    ABSENT = {"diffusers", "accelerate", "xformers", "cv2", "kornia", "torchsde", "torchdiffeq", "easing_functions",
              "generation_pb2", "generation_pb2_grpc", "engines_pb2", "engines_pb2_grpc", "tensors_pb2", "dashboard_pb2",
              "dashboard_pb2_grpc", "basicsr", "timm", "mmcv", "mmdet", "mmpose", "mmseg", "omegaconf", "pytorch_lightning",
              "colorama", "clip", "open_clip", "lycoris", "tensorizer", "resize_right", "gyre"}

    def __init__(self):
        self.made = []

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] not in self.ABSENT:
            return None
        import importlib.machinery
        self.made.append(fullname)
        return importlib.machinery.ModuleSpec(fullname, self, is_package=True)

    def create_module(self, spec):
        m = _AutoStubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def gyre_unified_pipeline():
    """Returns the reference's gyre.pipeline.unified_pipeline module (see the comment above)."""
    name = "gyre.pipeline.unified_pipeline"
    if name in sys.modules:
        return sys.modules[name]
    gyre_hires()               # real ResizeRight, easing stand-ins, gyre.pipeline.unet.{hires_fix, graft}
    k_diffusion()
    gyre_ddim()
    # hand-written stand-ins made earlier in this process are plain modules: let them answer for names nobody defined
    for mname, mod in list(sys.modules.items()):
        if mname.split(".")[0] in ("diffusers", "xformers", "easing_functions", "torchsde", "torchdiffeq") and \
                getattr(mod, "__file__", None) is None and not isinstance(mod, _AutoStubModule):
            if "__getattr__" not in mod.__dict__:
                mod.__getattr__ = (lambda mn: (lambda attr: _auto_attr(mn, attr)))(mname)
            if not hasattr(mod, "__path__"):
                mod.__path__ = []
    for pkg, rel in (("gyre", "gyre"), ("gyre.pipeline", "gyre/pipeline"), ("gyre.pipeline.unet", "gyre/pipeline/unet"),
                     ("gyre.pipeline.text_embedding", "gyre/pipeline/text_embedding"),
                     ("gyre.pipeline.schedulers", "gyre/pipeline/schedulers"),
                     ("gyre.pipeline.kschedulers", "gyre/pipeline/kschedulers"),
                     ("gyre.pipeline.controlnet", "gyre/pipeline/controlnet"),
                     ("gyre.pipeline.t2i_adapter", "gyre/pipeline/t2i_adapter"),
                     ("gyre.pipeline.models", "gyre/pipeline/models"), ("gyre.src", "gyre/src")):
        if pkg not in sys.modules:
            m = _AutoStubModule(pkg) if pkg.count(".") >= 2 else types.ModuleType(pkg)
            m.__path__ = [os.path.join(REF, rel)]
            sys.modules[pkg] = m
            parent, _, child = pkg.rpartition(".")
            if parent:
                setattr(sys.modules[parent], child, m)
    # gyre modules that are server / bookkeeping code (protobufs, logging UI, LoRA loaders, CLIP guidance ...): never needed
    # by the classes under test, expensive or impossible to import - stood in wholesale
    for mname in ("gyre.logging", "gyre.hints", "gyre.cache", "gyre.generated", "gyre.pipeline.lora",
                  "gyre.pipeline.lycoris", "gyre.pipeline.textual_inversion", "gyre.pipeline.latent_debugger",
                  "gyre.pipeline.vae_approximator", "gyre.pipeline.xformers_utils", "gyre.pipeline.unet.clipguided",
                  "gyre.pipeline.diffusers_types", "gyre.pipeline.attention_replacer", "gyre.pipeline.model_utils"):
        if mname not in sys.modules:
            m = _AutoStubModule(mname)
            m.__path__ = []
            sys.modules[mname] = m
            parent, _, child = mname.rpartition(".")
            setattr(sys.modules[parent], child, m)
    import transformers.models.clip as _tclip
    if not hasattr(_tclip, "CLIPFeatureExtractor"):           # removed alias of CLIPImageProcessor (transformers >= 5)
        _tclip.CLIPFeatureExtractor = _tclip.CLIPImageProcessor
    finder = _AutoStubFinder()
    sys.meta_path.append(finder)
    try:
        for _ in range(40):
            before = set(sys.modules)
            try:
                mod = importlib.import_module(name)
                break
            except ModuleNotFoundError as e:
                # one more absent third-party package: stand it in and start over (half-imported gyre modules dropped)
                missing = (e.name or "").split(".")[0]
                if not missing or missing in finder.ABSENT:
                    raise
                finder.ABSENT = set(finder.ABSENT) | {missing}
                for k in set(sys.modules) - before:
                    if k.startswith("gyre."):
                        del sys.modules[k]
        else:
            raise ImportError("gyre.pipeline.unified_pipeline: too many absent packages")
    finally:
        sys.meta_path.remove(finder)
    mod._auto_stubbed = sorted(set(finder.made))
    return mod
