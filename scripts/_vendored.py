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
    spec = importlib.util.spec_from_file_location("_gyre_sample_dpmpp_2m", p)
    mod = importlib.util.module_from_spec(spec)
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
    spec = importlib.util.spec_from_file_location("_gyre_scheduling_ddim", p)
    mod = importlib.util.module_from_spec(spec)
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
