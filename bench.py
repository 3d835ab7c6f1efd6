#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric: images/sec, SD1.5 txt2img 512x512, 50 Euler-a steps,
batch 8 per GPU (configs[1]); one "step" = one full pass of the hot path over one batch (50 CFG-doubled UNet
forwards + fused scheduler steps + VAE decode + image tail).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  Synthetic data: seeded random weights of the SD1.5 architecture and random
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

STEPS_PER_IMAGE = 50
H = W = 512
GUIDANCE = 7.5
# Algorithmic work (SURVEY.md 8d / BASELINE.md section 3): GFLOP
UNET_GFLOP_PER_SAMPLE_FWD = 803.3
VAE_DECODE_GFLOP = 2514.5
GFLOP_PER_IMAGE = 2 * STEPS_PER_IMAGE * UNET_GFLOP_PER_SAMPLE_FWD + VAE_DECODE_GFLOP   # 82 845


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
def cpu_reference(steps: int, warmup: int):
    """The reference's CPU path for this workload = the oracle (fp32 PyTorch restatement of the diffusers op
    graph + vendored sampler maths), timed on the host cores on a BOUNDED sample: per step one CFG-doubled UNet
    forward for ONE image at 64x64 latents; one VAE decode.  images/sec is extrapolated as
    1 / (50 * t_unet_step + t_decode)."""
    from oracle.unet import UNetConfig, OracleUNet, synth_params, unet_param_shapes
    from oracle.vae import VAEConfig, vae_decode, vae_param_shapes
    from oracle import sampling as osamp
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cfg = UNetConfig.sd15()
    P = synth_params(unet_param_shapes(cfg), seed=1234)
    unet = OracleUNet(cfg, P)
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(1, 77, 768, generator=g)
    unc = torch.randn(1, 77, 768, generator=g)
    cfgu = osamp.CFGParallel(unet, unc, emb, GUIDANCE)
    x = torch.randn(1, 4, 64, 64, generator=g)
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
        vae_decode(VP, vcfg, x / 0.18215)
        t_dec = time.perf_counter() - t0
    t_step = sum(ts) / len(ts)
    per_image = STEPS_PER_IMAGE * t_step + t_dec
    return {"value": 1.0 / per_image, "unit": "images/sec", "cores": cores, "kind": "port",
            "sample": f"1 image: {len(ts)} CFG-doubled UNet forwards (batch 2, 64x64 latents, fp32) timed at "
                      f"{t_step:.2f} s each + 1 VAE decode {t_dec:.2f} s; extrapolated to 50 steps + decode",
            "t_unet_step_s": t_step, "t_decode_s": t_dec}


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "SD1.5 txt2img 512x512, 50 Euler-a steps, CFG 7.5, batch 8 per GPU (BASELINE configs[1])",
              "per_gpu_batch": a.batch, "global_batch": a.batch * world, "steps_per_image": STEPS_PER_IMAGE,
              "parallelism": f"dp{world} (independent images sharded across GPUs; weights broadcast once; images gathered)",
              "l2": "no flush: per-step working set (1.9 GB weights + GBs of activations) exceeds the 126 MB L2"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        warm = max(1, min(a.warmup, 2))
        cb = cpu_reference(max(1, a.steps), warm)
        line = {"impl": "reference", "metric": "images/sec SD1.5 512x512 50-step", "value": cb["value"],
                "unit": "images/sec", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1000.0 * (STEPS_PER_IMAGE * cb["t_unet_step_s"] + cb["t_decode_s"]),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
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
    ucfg, vcfg = UNetConfig.sd15(), VAEConfig.sd()
    ushapes, vshapes = unet_param_shapes(ucfg), vae_param_shapes(vcfg)
    usd = synth_state_dict(ushapes, 1234, dtype=torch.float16, device=dev) if rank == 0 else None
    vsd = synth_state_dict(vshapes, 4321, dtype=torch.float16, device=dev) if rank == 0 else None
    usd = gdist.broadcast_state_dict(usd, ushapes, 0, dev)
    vsd = gdist.broadcast_state_dict(vsd, vshapes, 0, dev)
    unet = B200UNet(ucfg, dev).load_state_dict(usd)
    vae = B200VAE(vcfg, dev).load_state_dict(vsd)
    del usd, vsd
    pipe = B200Pipeline(unet, vae)

    B = a.batch
    g = torch.Generator().manual_seed(1000 + rank)
    emb_host = torch.randn(B, 77, 768, generator=g).half().pin_memory()
    unc_host = torch.randn(1, 77, 768, generator=g).half().expand(B, -1, -1).contiguous().pin_memory()
    emb_dev, unc_dev = emb_host.to(dev), unc_host.to(dev)
    img_host = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
    seed_base = 420420420 + rank * B

    def gens(step):
        # the reference builds its generators on the execution device (pipeline_wrapper.py:243-253)
        return [torch.Generator(dev).manual_seed(seed_base + 7919 * step + i) for i in range(B)]

    def run(step, e, u):
        out = pipe(e, u, height=H, width=W, num_inference_steps=STEPS_PER_IMAGE, guidance_scale=GUIDANCE,
                   generator=gens(step), sampler="k_euler_ancestral", output_type="uint8")
        return gdist.gather_images(out.images)

    def step_resident(step):
        return run(step, emb_dev, unc_dev)

    def step_e2e(step):
        e = emb_host.to(dev, non_blocking=True)
        u = unc_host.to(dev, non_blocking=True)
        out = pipe(e, u, height=H, width=W, num_inference_steps=STEPS_PER_IMAGE, guidance_scale=GUIDANCE,
                   generator=gens(step), sampler="k_euler_ancestral", output_type="uint8")
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
        N.prof_reset()
        N.prof_enable(True)
        step_resident(300)
        families = N.prof_read()
        N.prof_enable(False)
        N.prof_reset()
        peaks = measured_peaks()
        tc = {k: v for k, v in families.items() if v["flops"] > 0 and v["ms"] > 0}
        dom = max(tc, key=lambda k: tc[k]["ms"])
        ach = tc[dom]["flops"] / (tc[dom]["ms"] * 1e-3) / 1e12
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if dom in tj:
                # dram__bytes_read.sum + dram__bytes_write.sum per launch of this family from the committed ncu launch
                # list of the same workload (scripts/gpu_profile.sh), weighted like one step (50 forwards + 1 decode)
                traffic = tj[dom]["dram_bytes_per_launch"]
                traffic_src = "profiles/r01_traffic.json (ncu, cold L2 per launch)"
        roofline = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": tc[dom]["bytes"] / tc[dom]["count"],
                    "peak_source": peaks["source"] +
                    ", sustained figure (kernel timed inside a long step)",
                    "launches": tc[dom]["count"], "avg_launch_ms": tc[dom]["ms"] / tc[dom]["count"],
                    "algorithmic_gflop_per_launch": tc[dom]["flops"] / tc[dom]["count"] / 1e9,
                    "share_of_step_device_time": tc[dom]["ms"] / sum(v["ms"] for v in families.values())}
        for k, v in families.items():
            if v["ms"] > 0:
                v["tflops"] = v["flops"] / (v["ms"] * 1e-3) / 1e12
                v["gbs"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    line = {"metric": "images/sec SD1.5 512x512 50-step", "value": value, "unit": "images/sec", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate)", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e_value, "unit": "images/sec", "ms_per_step": ms_e2e / a.steps,
                    "h2d_bytes_per_step": int(emb_host.numel() * 2 + unc_host.numel() * 2),
                    "d2h_bytes_per_step": int(img_host.numel())},
            "gpu_launches": int(launches), "clocks": clk,
            "achieved_tflops_whole_step": GFLOP_PER_IMAGE * images / (ms / 1000.0) / 1000.0 / world}
    if roofline:
        line["roofline"] = roofline
        line["families"] = families
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(2, 1)
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
