#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric: images/sec, SD1.5 txt2img 512x512, 50 Euler-a steps,
batch 8 per GPU (configs[1]); one "step" = one full pass of the hot path over one batch (50 CFG-doubled UNet
forwards + fused scheduler steps + VAE decode + image tail).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference|torchlib]
                    [--config c2|c3|c4|c5] [--scaling weak|strong] [--batch B] [--cuda-graph]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

--config   c2 (default, the configuration the metric is quoted on) or one of the other BASELINE.json configurations:
           c3 SD1.5-inpaint 768x768, 64 Euler-a steps, two VAE encodes + decode, batch 4;
           c4 SD2.1-768-v, 50 Euler-a steps, ToMe r = N/2, global batch 16 (8 per GPU in weak mode);
           c5 SDXL-base topology 1024x1024, 30 steps, global batch 64 (8 per GPU in weak mode).
--scaling  weak (default): --batch images per GPU whatever N is.  strong: the configuration's GLOBAL batch (c2 8, c3 4,
           c4 16, c5 64, or --batch) is split over the N ranks, so per-GPU batches of 4 / 2 / 1 are what is timed.
--impl     b200 (the product) | reference (the oracle on the host cores: the reference's CPU path, a reported baseline)
           | torchlib (NOT the product: the same oracle graph in fp16 on this GPU through PyTorch's library kernels -
           cuDNN convolutions, cuBLASLt GEMMs, SDPA attention - i.e. what the reference's diffusers path becomes on
           current dependencies; the same-box comparator of SURVEY 8d).
--cuda-graph  whole-loop CUDA graph (gyre_b200.common_scheduler._loop_graphed).

Prints ONE JSON line (rank 0).  Synthetic data: seeded random weights of the named architecture and random
text embeddings (no checkpoints / tokenizer on the box); the arithmetic volume is identical to a real run.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GUIDANCE = 7.5
# Algorithmic work (SURVEY.md 8d / BASELINE.md section 3), GFLOP per image = 2 * steps * UNet sample-forward + VAE
CONFIGS = {
    "c2": {"workload": "SD1.5 txt2img 512x512, 50 Euler-a steps, CFG 7.5 (BASELINE configs[1])",
           "metric": "images/sec SD1.5 512x512 50-step", "unet": "sd15", "hw": 512, "steps": 50, "batch": 8,
           "global_batch": 8, "gflop_per_image": 2 * 50 * 803.3 + 2514.5, "ctx": 768},
    "c3": {"workload": "SD1.5-inpaint 768x768, 64 Euler-a steps, strength 1.0, two VAE encodes + decode, CFG 7.5 "
                       "(BASELINE configs[2]; hires_fix / grafted_inpaint off)",
           "metric": "images/sec SD1.5-inpaint 768x768 64-step", "unet": "sd15_inpaint", "hw": 768, "steps": 64, "batch": 4,
           "global_batch": 4, "gflop_per_image": 281363.0, "ctx": 768, "inpaint": True},
    "c4": {"workload": "SD2.1-768-v txt2img 768x768, 50 Euler-a steps, ToMe r = N/2 per block, CFG 7.5 (BASELINE configs[3])",
           "metric": "images/sec SD2.1 768x768 50-step ToMe 0.5", "unet": "sd21_v", "hw": 768, "steps": 50, "batch": 8,
           "global_batch": 16, "gflop_per_image": 197414.0, "ctx": 1024, "tome": 96 * 96 // 2},
    "c5": {"workload": "SDXL-base topology txt2img 1024x1024, 30 Euler-a steps, CFG 7.5 (BASELINE configs[4]; no reference "
                       "path exists for SDXL)",
           "metric": "images/sec SDXL-base 1024x1024 30-step", "unet": "sdxl", "hw": 1024, "steps": 30, "batch": 8,
           "global_batch": 64, "gflop_per_image": 416142.0, "ctx": 2048, "sdxl": True},
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        busy = [s for s, p in zip(sm, pw) if p > 300] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- CPU reference
def _oracle_unet_cfg(name):
    from oracle.unet import UNetConfig
    return {"sd15": UNetConfig.sd15, "sd15_inpaint": UNetConfig.sd15_inpaint, "sd21_v": UNetConfig.sd21_v,
            "sdxl": UNetConfig.sdxl}[name]()


def cpu_reference(cfgd: dict, steps: int, warmup: int):
    """The reference's CPU path for this workload = the oracle (fp32 PyTorch restatement of the diffusers op
    graph + vendored sampler maths), timed on the host cores on a BOUNDED sample: per step one CFG-doubled UNet
    forward for ONE image at the configuration's latent size; one VAE decode.  images/sec is EXTRAPOLATED as
    1 / (steps_per_image * t_unet_step + t_decode) - said so in `kind`."""
    from oracle.unet import OracleUNet, synth_params, unet_param_shapes
    from oracle.vae import VAEConfig, vae_decode, vae_param_shapes
    from oracle import sampling as osamp
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cfg = _oracle_unet_cfg(cfgd["unet"])
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    unet = OracleUNet(cfg, P)
    g = torch.Generator().manual_seed(11)
    lat = cfgd["hw"] // 8
    emb = torch.randn(1, 77, cfgd["ctx"], generator=g)
    unc = torch.randn(1, 77, cfgd["ctx"], generator=g)
    added = None
    if cfgd.get("sdxl"):
        added = {"text_embeds": torch.randn(2, 1280, generator=g), "time_ids": torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * 2)}
    cfgu = osamp.CFGParallel(unet, unc, emb, GUIDANCE, added)
    x = torch.randn(1, cfg.in_channels, lat, lat, generator=g)
    ts = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            cfgu(x, torch.tensor([981 - i]))
            dt = time.perf_counter() - t0
            if i >= warmup:
                ts.append(dt)
        vcfg = VAEConfig.sd()
        VP = synth_params(vae_param_shapes(vcfg, encoder=False), seed=4321)
        t0 = time.perf_counter()
        vae_decode(VP, vcfg, x[:, :4] / 0.18215)
        t_dec = time.perf_counter() - t0
    t_step = sum(ts) / len(ts)
    per_image = cfgd["steps"] * t_step + t_dec
    return {"value": 1.0 / per_image, "unit": "images/sec", "cores": cores, "kind": "port (extrapolated from a bounded sample)",
            "sample": f"1 image: {len(ts)} CFG-doubled UNet forwards (batch 2, {lat}x{lat} latents, fp32) timed at "
                      f"{t_step:.2f} s each + 1 VAE decode {t_dec:.2f} s; extrapolated to {cfgd['steps']} steps + decode",
            "t_unet_step_s": t_step, "t_decode_s": t_dec}


# ---------------------------------------------------------------------------------------------- torch library comparator
def torchlib_arm(cfgd: dict, B: int, steps: int, warmup: int, dev):
    """NOT the product.  The oracle's op graph evaluated in fp16 on this GPU by PyTorch's own library kernels (cuDNN
    convolutions, cuBLASLt GEMMs, SDPA attention, eager elementwise kernels): Euler-ancestral loop with parallel CFG +
    VAE decode + clamp, same shapes / step count / batch as the b200 arm.  It answers "what would the reference's
    diffusers path do on this box with current libraries" (second comparator)."""
    import oracle.unet as ou
    from oracle.unet import synth_params, unet_forward, unet_param_shapes
    from oracle.vae import VAEConfig, vae_decode, vae_param_shapes
    from oracle import sampling as osamp
    if cfgd.get("inpaint") or cfgd.get("tome"):
        raise SystemExit("--impl torchlib covers the txt2img configurations without ToMe (c2, c5)")
    ou.ATTENTION_IMPL = "sdpa"
    torch.backends.cudnn.benchmark = True
    cfg = _oracle_unet_cfg(cfgd["unet"])
    P = {k: v.to(dev).half() for k, v in synth_params(unet_param_shapes(cfg), seed=1234).items()}
    vcfg = VAEConfig.sd()
    VP = {k: v.to(dev).half() for k, v in synth_params(vae_param_shapes(vcfg, encoder=False), seed=4321).items()}
    g = torch.Generator().manual_seed(1000)
    lat = cfgd["hw"] // 8
    emb2 = torch.randn(2 * B, 77, cfgd["ctx"], generator=g).half().to(dev)
    added = None
    if cfgd.get("sdxl"):
        added = {"text_embeds": torch.randn(2 * B, 1280, generator=g).half().to(dev),
                 "time_ids": torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * (2 * B)).half().to(dev)}
    den = osamp.EpsDenoiser(lambda x, t: x, osamp.sd_alphas_cumprod())
    sig = osamp.k_sigmas(den, cfgd["steps"]).half().float()
    tsteps = den.sigma_to_t(sig[:-1]).to(dev)

    @torch.no_grad()
    def one_batch(seed):
        gen = torch.Generator(dev).manual_seed(seed)
        x = torch.randn(B, 4, lat, lat, generator=gen, device=dev, dtype=torch.float16) * sig[0]
        for i in range(cfgd["steps"]):
            s, s_next = sig[i], sig[i + 1]
            c_in = 1 / (s ** 2 + 1) ** 0.5
            x2 = torch.cat([x, x]) * c_in
            kw = {"added_cond_kwargs": added} if added is not None else {}
            eps = unet_forward(P, cfg, x2, tsteps[i].expand(2 * B), emb2, **kw)
            u, c = eps.chunk(2)
            e = u + GUIDANCE * (c - u)
            denoised = x - e * s
            sd, su = osamp.get_ancestral_step(s, s_next)
            d = (x - denoised) / s
            x = x + d * (sd - s)
            if s_next > 0:
                x = x + torch.randn(x.shape, generator=gen, device=dev, dtype=x.dtype) * su
        img = vae_decode(VP, vcfg, x / 0.18215)
        return ((img / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8)

    for i in range(max(1, warmup)):
        one_batch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one_batch(100 + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torchlib"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--batch", type=int, default=None, help="weak: images per GPU per step; strong: global batch")
    ap.add_argument("--cuda-graph", action="store_true")
    ap.add_argument("--hires-fix", action="store_true",
                    help="c3 as the reference runs it by default: natural-size twin + cross-blend while u < 0.667")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    a = ap.parse_args()
    cfgd = CONFIGS[a.config]
    STEPS_PER_IMAGE, H, W = cfgd["steps"], cfgd["hw"], cfgd["hw"]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.scaling == "weak":
        B = a.batch or cfgd["batch"]
        global_batch = B * world
    else:
        global_batch = a.batch or cfgd["global_batch"]
        if global_batch % world != 0:
            raise SystemExit(f"strong scaling: global batch {global_batch} does not split over {world} ranks")
        B = global_batch // world
    config = {"workload": cfgd["workload"] + f", batch {B} per GPU", "config": a.config,
              "per_gpu_batch": B, "global_batch": global_batch, "steps_per_image": STEPS_PER_IMAGE,
              "parallelism": f"dp{world} (independent images sharded across GPUs; weights broadcast once; images gathered)",
              "cuda_graph": bool(a.cuda_graph), "hires_fix": bool(a.hires_fix),
              "l2": "no flush: per-step working set (GBs of weights + activations) exceeds the 126 MB L2"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        warm = max(1, min(a.warmup, 2))
        cb = cpu_reference(cfgd, max(1, a.steps), warm)
        line = {"impl": "reference", "metric": cfgd["metric"], "value": cb["value"],
                "unit": "images/sec", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1000.0 * (STEPS_PER_IMAGE * cb["t_unet_step_s"] + cb["t_decode_s"]),
                "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    if not torch.cuda.is_available():
        raise SystemExit(f"bench.py --impl {a.impl} needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    if a.impl == "torchlib":
        if rank != 0:
            return 0
        ms = torchlib_arm(cfgd, B, a.steps, a.warmup, dev)
        images = B * a.steps
        line = {"impl": "torchlib", "note": "NOT the product: the oracle graph in fp16 through PyTorch library kernels (cuDNN / "
                "cuBLASLt / SDPA) on the same GPU", "metric": cfgd["metric"], "value": images / (ms / 1000.0),
                "unit": "images/sec", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
                "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": config, "gpu_launches": 0, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
        print(json.dumps(line))
        return 0

    import torch.distributed as dist
    json_fd = None
    if world > 1:
        # NCCL prints its banner ("NCCL version ...") with printf on the process's stdout whatever NCCL_DEBUG_FILE says,
        # and the contract is ONE JSON line there: fd 1 is pointed at stderr for the whole run and the JSON line goes
        # out through a duplicate of the original stdout
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    from gyre_b200 import _native as N
    from gyre_b200 import dist as gdist
    from gyre_b200.config import UNetConfig, VAEConfig
    from gyre_b200.pipeline import B200Pipeline
    from gyre_b200.unet import B200UNet
    from gyre_b200.vae import B200VAE
    from gyre_b200.weights import synth_state_dict, unet_param_shapes, vae_param_shapes
    N.load()

    # ---- weights: created on rank 0, ONE NCCL broadcast to the other ranks
    ucfg = {"sd15": UNetConfig.sd15, "sd15_inpaint": UNetConfig.sd15_inpaint, "sd21_v": UNetConfig.sd21_v,
            "sdxl": UNetConfig.sdxl}[cfgd["unet"]]()
    vcfg = VAEConfig.sd()
    ushapes, vshapes = unet_param_shapes(ucfg), vae_param_shapes(vcfg)
    usd = synth_state_dict(ushapes, 1234, dtype=torch.float16, device=dev) if rank == 0 else None
    vsd = synth_state_dict(vshapes, 4321, dtype=torch.float16, device=dev) if rank == 0 else None
    usd = gdist.broadcast_state_dict(usd, ushapes, 0, dev)
    vsd = gdist.broadcast_state_dict(vsd, vshapes, 0, dev)
    unet = B200UNet(ucfg, dev, hold_parameters=False).load_state_dict(usd)
    vae = B200VAE(vcfg, dev, hold_parameters=False).load_state_dict(vsd)
    del usd, vsd
    pipe = B200Pipeline(unet, vae)
    pipe.use_cuda_graph = bool(a.cuda_graph)
    if cfgd.get("tome"):
        pipe.set_options({"tome": cfgd["tome"]})

    g = torch.Generator().manual_seed(1000 + rank)
    ctx = cfgd["ctx"]
    emb_host = torch.randn(B, 77, ctx, generator=g).half().pin_memory()
    unc_host = torch.randn(1, 77, ctx, generator=g).half().expand(B, -1, -1).contiguous().pin_memory()
    emb_dev, unc_dev = emb_host.to(dev), unc_host.to(dev)
    img_host = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
    seed_base = 420420420 + rank * B
    extra_host, extra_dev = {}, {}
    if cfgd.get("inpaint"):
        im = torch.rand(1, 3, H, W, generator=g)
        mk = torch.zeros(1, 1, H, W)
        mk[:, :, H // 4:3 * H // 4, W // 3:5 * W // 6] = 1
        extra_host = {"image": im.pin_memory(), "mask_image": mk.pin_memory()}
        extra_dev = {k: v.to(dev) for k, v in extra_host.items()}
    if cfgd.get("sdxl"):
        extra_host = {"added_cond_kwargs": {"text_embeds": torch.randn(B, 1280, generator=g).pin_memory(),
                                            "time_ids": torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * B).pin_memory()}}
        extra_dev = {"added_cond_kwargs": {k: v.to(dev) for k, v in extra_host["added_cond_kwargs"].items()}}
    fixed = {"strength": 1.0} if cfgd.get("inpaint") else {}
    fixed["hires_fix"] = bool(a.hires_fix)      # BASELINE configs state hires_fix=False; only c3 is above native size
    h2d_extra = sum(v.numel() * v.element_size() for v in
                    (extra_host.get("added_cond_kwargs", {}).values() if cfgd.get("sdxl") else extra_host.values()))

    def gens(step):
        # the reference builds its generators on the execution device (pipeline_wrapper.py:243-253)
        return [torch.Generator(dev).manual_seed(seed_base + 7919 * step + i) for i in range(B)]

    def to_dev(d):
        return {k: ({kk: vv.to(dev, non_blocking=True) for kk, vv in v.items()} if isinstance(v, dict)
                    else v.to(dev, non_blocking=True)) for k, v in d.items()}

    def call(step, e, u, extra):
        return pipe(e, u, height=H, width=W, num_inference_steps=STEPS_PER_IMAGE, guidance_scale=GUIDANCE,
                    generator=gens(step), sampler="k_euler_ancestral", output_type="uint8", **extra, **fixed)

    def step_resident(step):
        return gdist.gather_images(call(step, emb_dev, unc_dev, extra_dev).images)

    def step_e2e(step):
        e = emb_host.to(dev, non_blocking=True)
        u = unc_host.to(dev, non_blocking=True)
        out = call(step, e, u, to_dev(extra_host))
        img_host.copy_(out.images, non_blocking=True)          # device -> pinned host read of the result
        gdist.gather_images(out.images)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, first_step):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(k):
            fn(first_step + s)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for s in range(a.warmup):
        step_resident(s)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = N.launch_count()
    ms = timed(step_resident, a.steps, 100)
    launches = N.launch_count() - l0
    ms_e2e = timed(step_e2e, a.steps, 200)
    clk = clocks.stop() if rank == 0 else None

    images = B * world * a.steps
    value = images / (ms / 1000.0)
    e2e_value = images / (ms_e2e / 1000.0)

    # ---- roofline of the dominant kernel family, measured live with CUDA events on the launching stream
    roofline, families = None, None
    if not a.no_profile:
        was_graph = pipe.use_cuda_graph
        pipe.use_cuda_graph = False                 # the per-family events bracket individual launches
        N.prof_reset()
        N.prof_enable(True)
        step_resident(300)
        peaks = measured_peaks()
        families = N.prof_read(peaks["tflops_sustained"], peaks["hbm_gbs"])
        N.prof_enable(False)
        N.prof_reset()
        pipe.use_cuda_graph = was_graph
        tc = {k: v for k, v in families.items() if v["flops"] > 0 and v["ms"] > 0}
        dom = max(tc, key=lambda k: tc[k]["ms"])
        ach = tc[dom]["flops"] / (tc[dom]["ms"] * 1e-3) / 1e12
        traffic, traffic_src = None, None
        for tname in ("r02_traffic.json", "r01_traffic.json"):
            tp = os.path.join(ROOT, "profiles", tname)
            if a.config == "c2" and os.path.exists(tp):
                tj = json.load(open(tp))
                if dom in tj:
                    # dram__bytes_read.sum + dram__bytes_write.sum per launch of this family from the committed ncu launch
                    # list of the same workload (scripts/gpu_profile.sh), weighted like one step (50 forwards + 1 decode)
                    traffic = tj[dom]["dram_bytes_per_launch"]
                    traffic_src = f"profiles/{tname} (ncu, cold L2 per launch)"
                    break
        roofline = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": tc[dom]["bytes"] / tc[dom]["count"],
                    "peak_source": peaks["source"] +
                    ", sustained figure (kernel timed inside a long step)",
                    "launches": tc[dom]["count"], "avg_launch_ms": tc[dom]["ms"] / tc[dom]["count"],
                    "algorithmic_gflop_per_launch": tc[dom]["flops"] / tc[dom]["count"] / 1e9,
                    "share_of_step_device_time": tc[dom]["ms"] / sum(v["ms"] for v in families.values()),
                    # the family mixes tensor-bound and HBM-bound shapes (K = 320 GEMMs are HBM-bound): the sum over its
                    # launches of max(flops / tensor peak, algorithmic bytes / HBM peak) against the measured time
                    "frac_vs_binding_roof_per_launch": tc[dom]["roofline_ms"] / tc[dom]["ms"],
                    "hbm_peak": peaks["hbm_gbs"]}
        for k, v in families.items():
            if v["ms"] > 0:
                v["tflops"] = v["flops"] / (v["ms"] * 1e-3) / 1e12
                v["gbs"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
                v["frac_vs_binding_roof"] = v["roofline_ms"] / v["ms"]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    line = {"metric": cfgd["metric"], "value": value, "unit": "images/sec", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": a.scaling, "vs_baseline": None, "dtype": "f16 (fp32 accumulate)", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e_value, "unit": "images/sec", "ms_per_step": ms_e2e / a.steps,
                    "h2d_bytes_per_step": int(emb_host.numel() * 2 + unc_host.numel() * 2 + h2d_extra),
                    "d2h_bytes_per_step": int(img_host.numel())},
            "gpu_launches": int(launches), "clocks": clk,
            "achieved_tflops_whole_step": cfgd["gflop_per_image"] * images / (ms / 1000.0) / 1000.0 / world}
    if roofline:
        line["roofline"] = roofline
        line["families"] = families
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(cfgd, 2, 1)
    if world > 1:
        dist.destroy_process_group()
    if json_fd is not None:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
