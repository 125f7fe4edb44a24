#!/usr/bin/env python
"""Headline benchmark of the stamp path (BASELINE.json: stamps/sec & p50 ms/stamp, 512x512 20-step image-conditioned
inpaint at 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W          # N > 1: launched by torchrun, one rank per GPU
  python bench.py --impl reference ...                   # the reference's CPU path (oracle port) on the host cores

A "step" is one stamp per GPU: canvas pre-process -> 2x VAE encode -> 20 three-branch UNet evaluations + guidance/DDIM
-> VAE decode -> composite, on synthetic inputs (seeded weights in the diffusers key inventory; no checkpoints offline).
`value` = stamps/s with the canvas resident in HBM (CUDA events, max over ranks); `e2e` = the same through the
handler-facing call with pinned HOST uint8 buffers, H2D / D2H (and the NCCL scatter / gather for N > 1) inside the timed
region. One JSON line on stdout (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY.md Appendix C / BASELINE.md §3: algorithmic GFLOP per sample (2*MAC of conv / linear / attention matmuls)
GFLOP_UNET = {64: 11.2, 128: 43.3, 256: 177.0, 512: 798.1}
GFLOP_VAE_ENC = {64: 16.9, 128: 67.8, 256: 272.7, 512: 1116.7}
GFLOP_VAE_DEC = {64: 38.8, 128: 155.1, 256: 622.2, 512: 2514.5}


def stamp_flops(R, n_eval):
    return (3 * n_eval * GFLOP_UNET[R] + 2 * GFLOP_VAE_ENC[R] + GFLOP_VAE_DEC[R]) * 1e9


def attention_core_flops(R, n_eval):
    """Self-attention core (QK^T and PV, 4*seq^2*C per layer and sample; SURVEY.md Appendix C counting rules) of the UNet:
    5 / 5 / 5 / 1 transformer layers at 320 / 640 / 1280 / 1280 channels. These run in flash_attn_kernel, every other
    contraction (conv, linear, folded cross-attention, the 512-wide VAE attention) in gemm_tc_kernel."""
    h = R // 8
    per_sample = sum(n * 4.0 * (h * h / 4 ** lvl) ** 2 * c for lvl, (n, c) in enumerate([(5, 320), (5, 640), (5, 1280), (1, 1280)]))
    return 3 * n_eval * per_sample


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1400.0, 1590.0, "fallback"


def committed_traffic(R, S, B):
    """dram__bytes_read.sum + dram__bytes_write.sum of the contraction kernel, average per launch, from the committed ncu
    pass over one stamp of this workload (profiles/traffic_r1.json); None for other workloads."""
    p = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    if (d.get("resolution"), d.get("denoise_steps"), d.get("batch")) != (R, S, B):
        return None
    return d.get("avg_dram_bytes_per_launch")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(self.rows)}


def synthetic_inputs(R, n, seed=2):
    import torch
    from diffusiontexturepainting_b200.testdata import make_canvas, smooth_image
    brush = smooth_image(1, 3, R)
    canvases = make_canvas(n, R, seed)
    return brush, canvases


# ----------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (fp32 PyTorch port of the reference path) on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_unet_eval_seconds(R, reps, warm, branches=3):
    import torch
    from diffusiontexturepainting_b200 import weights as W
    from oracle import unet as un
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = W.sd15_config()
    sd = W.merge_lora(W.synth_state_dict(W.unet_param_shapes(cfg.unet), 20240726))
    g = torch.Generator().manual_seed(0)
    h = R // 8
    x = torch.randn(branches, 9, h, h, generator=g)
    ctx = torch.randn(branches, 14, 768, generator=g)
    times = []
    with torch.inference_mode():
        for i in range(warm + reps):
            t0 = time.perf_counter()
            un.unet_forward(sd, cfg.unet, x, 501.0, ctx)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    R, S = args.resolution, args.denoise_steps
    # bounded sample: one 3-branch UNet evaluation of the same 512x512 workload per bench step; a stamp is
    # F(R,S) / (3*U(R)) such evaluations of algorithmic work (the UNet loop is >= 90 % of a stamp's FLOPs)
    probe = cpu_unet_eval_seconds(R, 1, 0, branches=1)[0]
    branches = 3 if probe * 3 * (args.steps + args.warmup) < 240 else 1
    times = cpu_unet_eval_seconds(R, args.steps, args.warmup, branches=branches)
    t_eval = statistics.median(times) * (3.0 / branches)
    scale = stamp_flops(R, S) / (3 * GFLOP_UNET[R] * 1e9)
    t_stamp = t_eval * scale
    value = 1.0 / t_stamp
    cores = os.cpu_count() or 1
    sample = (f"{branches}-branch fp32 UNet evaluation at {R}x{R} per step (oracle port of the reference path, torch CPU, "
              f"{cores} threads), scaled x{scale:.2f} by algorithmic FLOPs to one {S}-evaluation stamp")
    line = {"impl": "reference", "metric": "stamps/sec", "value": value, "unit": "stamps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_stamp * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, 1),
            "cpu_baseline": {"value": value, "unit": "stamps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "stamps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def bench_config(args, world):
    return {"workload": f"C2: {args.resolution}x{args.resolution} stamp, {args.denoise_steps}-step DDIM "
                        f"({args.denoise_steps} UNet evaluations x 3 guidance branches, strict schedule), cfg 2.0, tg 1.0, "
                        f"context_pad 150, B={args.batch} stamp per GPU per step",
            "resolution": args.resolution, "denoise_steps": args.denoise_steps, "unet_evaluations": args.denoise_steps,
            "batch_per_gpu": args.batch, "parallelism": f"dp{world} (independent stamps per GPU, no data-path collective)",
            "l2": "inputs+weights (1.9 GB fp16) exceed L2 (126 MB); no flush needed",
            "weights": "synthetic seeded, diffusers key inventory, LoRA rank-4 merged"}


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from diffusiontexturepainting_b200 import parallel as par
    from diffusiontexturepainting_b200 import weights as W
    from diffusiontexturepainting_b200.trt_model import TRTConditionalInpainter

    rank, world, local = par.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    R, S, B = args.resolution, args.denoise_steps, args.batch
    cfg = W.sd15_config()

    # weights: generated on rank 0, packed, broadcast once over NCCL, loaded from device memory
    if world > 1:
        from diffusiontexturepainting_b200.engine import Engine
        packed = None
        uncond = torch.empty(1, 14, cfg.enc.cross_dim, device=dev)
        if rank == 0:
            u, v, e = W.synth_model(cfg)
            packed = {"unet." + k: t for k, t in W.pack_unet(W.merge_lora(u)).items()}
            packed.update({"vae." + k: t for k, t in W.pack_vae(v).items()})
            enc = W.pack_encoder(e)
            enc["pos_emb"] = W.patch_pos_emb(cfg.enc.width, cfg.enc.num_patches).reshape(-1, cfg.enc.width)
            packed.update({"enc." + k: t for k, t in enc.items()})
            uncond.copy_(e["uncond_vector"])
            del u, v, e
        packed = par.broadcast_packed(packed, dev, src=0)
        dist.broadcast(uncond, src=0)
        model = TRTConditionalInpainter(R, device=local, model_config=cfg, max_batch_size=B, preloaded=(packed, uncond))
        del packed
    else:
        model = TRTConditionalInpainter(R, device=local, model_config=cfg, max_batch_size=B)
    model.pipeline.strict_schedule = True   # S evaluations (the reference's steps=S would run S-1)
    model.pipeline.sample_posterior = True
    eng = model.engine

    brush, canvases = synthetic_inputs(R, B * world)
    model.set_brush(brush)
    settings = dict(steps=S, context_pad=150, tg_steps=S, width=R, cfg_weight=2.0, tg_weight=1.0)
    lo, hi = par.shard_range(B * world, rank, world)
    canvas_dev = canvases[lo:hi].to(dev)
    h = R // 8
    lat = torch.randn(B, 4, h, h, generator=torch.Generator().manual_seed(42 + rank)).to(dev)
    vn = torch.randn(2 * B, 4, h, h, generator=torch.Generator().manual_seed(7 + rank)).to(dev)
    out_f32 = torch.empty(B, 3, R, R, device=dev)
    model.pipeline.update_infer_settings(S, 2.0, 1.0, S)
    model.pipeline._push_schedule(1.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        eng.stamp(canvas_dev, model.image, 150, lat, vn, composite=True, out_f32=out_f32)

    # ---- device-resident throughput --------------------------------------------------------------------------
    for _ in range(args.warmup):
        resident_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.counter("launches")
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for a, b in evs:
        a.record()
        resident_step()
        b.record()
    t_all1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.counter("launches") - launches0
    total_ms = t_all0.elapsed_time(t_all1)
    per = sorted(a.elapsed_time(b) for a, b in evs)
    tmax = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    value = args.steps * B * world / (total_ms / 1e3)

    # ---- per-kernel-family device time of one stamp (CUDA events around every launch, same stream) -------------
    eng.set_option("profile", 1)
    resident_step()
    torch.cuda.synchronize()
    prof = eng.profile()
    eng.set_option("profile", 0)

    # ---- in-graph cost of each kernel family: stamp time with the family's launches left out of the captured graph
    # (results are garbage while a family is skipped; the difference to the full stamp is what the family costs in situ,
    # launch gaps and programmatic-launch overlap included) ---------------------------------------------------------
    in_graph = None
    if rank == 0 and not args.no_ablation:
        def timed(n=3):
            for _ in range(2):
                resident_step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                resident_step()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n
        base = timed()
        in_graph = {"stamp_ms": base}
        for k, name in ((1, "contraction"), (2, "groupnorm"), (3, "layernorm"), (6, "flash_attn")):
            eng.set_option("debug_skip_kinds", 1 << k)
            in_graph[name + "_ms"] = base - timed()
        eng.set_option("debug_skip_kinds", 0)
        resident_step()
        torch.cuda.synchronize()

    # ---- end to end through the handler-facing call: pinned host uint8 RGBA in, uint8 RGB out ------------------
    host_in = (canvases.permute(0, 2, 3, 1) * 255).to(torch.uint8).contiguous().pin_memory() if rank == 0 else None
    host_out = torch.empty(B * world, R, R, 3, dtype=torch.uint8).pin_memory() if rank == 0 else None

    def e2e_step():
        if world > 1:
            full = host_in.to(dev, non_blocking=True) if rank == 0 else None
            mine = par.scatter_stamps(full, (B, R, R, 4), torch.uint8, dev, src=0)
        else:
            mine = host_in.to(dev, non_blocking=True)
        res = model.stamp_u8(mine, init_latents=lat, vae_noise=vn, **settings)
        allres = par.gather_stamps(res, dst=0)
        if rank == 0:
            host_out.copy_(allres, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    n_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_e2e * B * world / float(te.item())

    if rank == 0:
        hbm, tf_sus, tf_burst, which = measured_peaks()
        gemm_us, gemm_n = prof["contraction"]
        flash_us, flash_n = prof.get("flash_attn", (0, 0))
        flops_all = stamp_flops(R, S) * B
        flops_flash = attention_core_flops(R, S) * B
        flops = flops_all - flops_flash  # algorithmic FLOPs executed by the contraction kernel
        achieved = flops / (gemm_us * 1e-6) / 1e12 if gemm_us else None
        total_prof = sum(v[0] for v in prof.values())
        cpu = None
        if not args.no_cpu_baseline:
            t = min(cpu_unet_eval_seconds(R, 2, 0, branches=3))
            scale = stamp_flops(R, S) / (3 * GFLOP_UNET[R] * 1e9)
            cpu = {"value": 1.0 / (t * scale), "unit": "stamps/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"two three-branch fp32 UNet evaluations at {R}x{R} (oracle port of the reference path, torch "
                             f"CPU, all host threads; best of 2), scaled x{scale:.2f} by algorithmic FLOPs to one "
                             f"{S}-evaluation stamp"}
        line = {
            "metric": "stamps/sec", "value": value, "unit": "stamps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "p50_ms_per_stamp": per[len(per) // 2] / B,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate)",
            "data": "synthetic", "config": bench_config(args, world),
            "e2e": {"value": e2e_value, "unit": "stamps/s", "h2d_bytes_per_step": B * world * R * R * 4,
                    "d2h_bytes_per_step": B * world * R * R * 3, "steps": n_e2e},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 implicit-GEMM conv / linear / attention)",
                         "achieved": achieved, "peak": tf_sus, "unit": "TFLOP/s",
                         "frac": achieved / tf_sus if achieved else None, "traffic": committed_traffic(R, S, B),
                         "peak_source": which,
                         "algorithmic_flops_in_kernel_per_stamp": flops, "algorithmic_flops_per_stamp": flops_all,
                         "whole_stamp_tflops_per_gpu": flops_all / B * value / world / 1e12,
                         "whole_stamp_frac_of_peak": flops_all / B * value / world / 1e12 / tf_sus,
                         "launches_per_stamp": gemm_n,
                         "avg_launch_us": gemm_us / gemm_n if gemm_n else None,
                         "share_of_step": gemm_us / total_prof if total_prof else None,
                         "achieved_in_graph": (flops / (in_graph["contraction_ms"] * 1e-3) / 1e12
                                               if in_graph and in_graph.get("contraction_ms", 0) > 0 else None),
                         "note": "achieved: CUDA events around every contraction launch of one stamp in eager order "
                                 "(includes ~2-3 us of launch gap per launch); achieved_in_graph: same FLOPs over the "
                                 "stamp-time difference with the contraction launches removed from the CUDA graph"},
            "roofline_flash_attn": {"bound": "tensor", "kernel": "flash_attn2_kernel / flash_attn_kernel (tcgen05)",
                                    "achieved": flops_flash / (flash_us * 1e-6) / 1e12 if flash_us else None,
                                    "peak": tf_sus, "unit": "TFLOP/s", "algorithmic_flops_per_stamp": flops_flash,
                                    "launches_per_stamp": flash_n,
                                    "note": "one MUFU ex2 per score: 16/clk/SM caps d=40 heads near 0.19 of the tensor peak"},
            "kernel_time_us_per_stamp": {k: v[0] for k, v in prof.items()},
            "kernel_launches_per_stamp": {k: v[1] for k, v in prof.items()},
            "kernel_time_in_graph_ms": in_graph,
            "cpu_baseline": cpu, "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything that libraries print to fd 1 (e.g. the NCCL version banner) goes to stderr; the single JSON line is
    written to the original stdout by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--resolution", type=int, default=512)
    ap.add_argument("--denoise-steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=1, help="stamps per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ablation", action="store_true", help="skip the in-graph per-family cost measurement")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
