#!/usr/bin/env python
"""Benchmark of the Multi-HMR hot path (BASELINE.json: images/sec, multiHMR_896_L, batch 8 per GPU).

  python bench.py --gpus N --steps K --warmup W            this repo: sm_100a engine through the C-ABI
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the reference's algorithm on the
                                                           host CPU cores (oracle port, fp32 PyTorch)

One "step" = one pass of `Model.forward(x, K)` over one batch of 8 synthetic 896x896 images per GPU
(random-init weights of the ViT-L architecture, seeded).  N > 1 is launched by torchrun, one rank per GPU,
image shards per rank (weak scaling) + one NCCL all-gather of the per-person records per step.

Prints ONE JSON line (rank 0):  value = whole-job images/s with inputs resident in HBM; e2e = the same
metric through the public API with pinned HOST inputs (H2D) and host outputs (D2H) inside the timed
region; roofline = live CUDA-event timing of the dominant kernel family vs the measured peak;
cpu_baseline = the oracle port timed on this box's host cores on a bounded sample (1 image).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="multiHMR_896_L", backbone="dinov2_vitl14", img_size=896, batch_per_gpu=8,
                det_thresh=0.3, nms_kernel_size=3, target_persons_per_image=2, seed=0)
ARCH = {"dinov2_vits14": (384, 12), "dinov2_vitb14": (768, 12), "dinov2_vitl14": (1024, 24)}
METRIC = "images/sec multiHMR_896_L bs=8"


def vit_flops_per_image(backbone: str, img_size: int) -> float:
    """depth*(24 T D^2 + 4 T^2 D) + 2 N 588 D (matmul 2mnk only)."""
    D, depth = ARCH[backbone]
    N = (img_size // 14) ** 2
    T = N + 1
    return depth * (24.0 * T * D * D + 4.0 * T * T * D) + 2.0 * N * 588 * D


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(tflops=float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))),
                    tflops_burst=float(p.get("bf16_tflops", 1590.0)), hbm_gbs=float(p.get("hbm_gbs", 6650.0)),
                    source="MEASURED_PEAKS.json (measured)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm_gbs=6650.0, source="B200_PROFILING.md fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md): the
    sampler runs from before the warm-up, and only samples stamped inside [mark_start, mark_stop] count."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.t0, self.t1 = index, None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def mark_start(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def mark_stop(self):
        import datetime
        self.t1 = datetime.datetime.now()

    def stop(self) -> dict:
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons, n_all = [], [], [], set(), 0
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            f = [c.strip() for c in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                vals = (float(f[1]), float(f[2]), float(f[3]))
            except ValueError:
                continue
            n_all += 1
            if self.t0 is not None and not (self.t0 <= ts <= self.t1):
                continue
            sm.append(vals[0]); mx.append(vals[1]); pw.append(vals[2])
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "samples_total": n_all,
                    "reasons": ["no samples inside the timed region"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
def build_workload(det_bias: float):
    from multihmr_b200 import synth

    w = WORKLOAD
    sd = synth.make_state_dict(w["backbone"], w["img_size"], seed=w["seed"], det_bias=det_bias)
    bm = synth.make_body_model(w["seed"])
    return sd, bm


def calibrate_det_bias(model, x, K, target_total: int, det_thresh: float) -> float:
    """Random-init weights have no meaningful detection density: shift the detection logit so that about
    `target_total` NMS maxima pass the threshold on this batch (setup, untimed)."""
    import torch

    res = model.res
    idx = (torch.zeros(1, dtype=torch.int64),) * 4
    out = model(x, idx=idx, K=K, is_training=True)  # training-style: raw sigmoid scores, no NMS
    s = out["scores"][..., 0].float().cpu().clamp(1e-4, 1 - 1e-4)
    logit = torch.log(s / (1 - s))
    mx = torch.nn.functional.max_pool2d(logit[:, None], 3, 1, 1)[:, 0]
    peaks = logit[(mx == logit)].flatten().sort(descending=True).values
    k = min(target_total, peaks.numel() - 1)
    cut = 0.5 * (peaks[k - 1] + peaks[k]).item()
    want = math.log(det_thresh / (1 - det_thresh))
    return want - cut  # added to the current bias (0)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from multihmr_b200 import parallel, synth
    from multihmr_b200.api import forward_model
    from multihmr_b200.model import Model

    world, rank, local = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) out of it
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOAD
    B, S = w["batch_per_gpu"], w["img_size"]
    max_persons = 128

    # ---- setup (untimed): weights, calibration of the synthetic detection density, final engine
    x_host = synth.make_images(B, S, seed=w["seed"] + rank).pin_memory()
    K_host = synth.make_cameras(B, S, seed=w["seed"] + rank).pin_memory()
    sd, bm = build_workload(det_bias=0.0)
    model = Model(backbone=w["backbone"], img_size=S, max_batch=B, max_persons=max_persons, body_model=bm, device=dev)
    model.load_state_dict(sd)
    x_dev, K_dev = x_host.to(dev), K_host.to(dev)
    shift = calibrate_det_bias(model, x_dev, K_dev, w["target_persons_per_image"] * B, w["det_thresh"])
    del model
    torch.cuda.empty_cache()
    sd["mlp_classif.2.bias"] = sd["mlp_classif.2.bias"] + shift
    model = Model(backbone=w["backbone"], img_size=S, max_batch=B, max_persons=max_persons, body_model=bm, device=dev)
    model.load_state_dict(sd)
    model.finalize()
    del sd

    def step_device():
        t, P = model.forward_raw(x_dev, K_dev, det_thresh=w["det_thresh"], nms_kernel_size=w["nms_kernel_size"])
        if world > 1:
            rec = parallel.pack_records(t, P, rank * B, max_persons, model.num_betas, model.num_verts)
            parallel.all_gather_persons(rec, P)
        return P

    host_out = {}

    def step_e2e():
        # public API with HOST buffers: pinned H2D of the images, forward, D2H of every person tensor
        persons = forward_model(model, x_host, K_host, det_thresh=w["det_thresh"], nms_kernel_size=w["nms_kernel_size"])
        t = model.last_outputs
        P = len(persons)
        nbytes = 0
        for k in ("det_score", "loc", "transl", "transl_pelvis", "rotvec", "expression", "shape", "v3d", "j3d", "j2d"):
            src = t[k][:P]
            if k not in host_out or host_out[k].shape[0] < P:
                host_out[k] = torch.empty((max_persons,) + tuple(src.shape[1:]), dtype=src.dtype).pin_memory()
            host_out[k][:P].copy_(src, non_blocking=True)
            nbytes += src.numel() * src.element_size()
        if world > 1:
            rec = parallel.pack_records(t, P, rank * B, max_persons, model.num_betas, model.num_verts)
            parallel.all_gather_persons(rec, P)
        torch.cuda.current_stream().synchronize()
        return P, nbytes

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.mark_start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(steps):
            last = fn()
        e1.record()
        torch.cuda.synchronize()
        if sampler:
            sampler.mark_stop()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), last, clocks

    # ---- device-resident throughput (value) with live per-kernel-family CUDA-event timing
    model.set_profiling(True)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    model.get_profile()  # drop warm-up records
    ms_total, P_last, clocks = timed(step_device, args.steps, 0, sampler)
    prof = model.get_profile()
    launches = model.last_launch_count()
    model.set_profiling(False)
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- end to end through the public API with host buffers
    ms_e2e, last, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    P_e2e, d2h_bytes = last
    h2d_bytes = x_host.numel() * 4 + K_host.numel() * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    D, depth = ARCH[w["backbone"]]
    N = (S // 14) ** 2
    T, M = N + 1, B * (N + 1)
    algo_flops = {  # algorithmic FLOPs per launch (2mnk)
        "gemm_qkv": 2.0 * M * 3 * D * D, "gemm_proj": 2.0 * M * D * D, "gemm_fc1": 2.0 * M * 4 * D * D,
        "gemm_fc2": 2.0 * M * 4 * D * D, "attention": 4.0 * B * T * T * D,
    }
    fam = {}
    for k, fl in algo_flops.items():
        ms, n = prof[k]
        if n:
            fam[k] = dict(ms_per_launch=ms / n, launches_per_step=n / args.steps,
                          tflops=fl / (ms / n) / 1e9, share_of_step=ms / ms_total)
    dom = max(fam, key=lambda k: fam[k]["share_of_step"]) if fam else None
    vit_ms = sum(prof[k][0] for k in ("misc", "layernorm", "gemm_qkv", "attention", "gemm_proj", "gemm_fc1", "gemm_fc2"))
    vit_tflops = vit_flops_per_image(w["backbone"], S) * B * args.steps / max(vit_ms, 1e-9) / 1e9
    roofline = None
    if dom:
        traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get(dom, {}).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": dom, "achieved": round(fam[dom]["tflops"], 1),
                    "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": round(fam[dom]["tflops"] / peaks["tflops"], 4),
                    "traffic": traffic, "peak_source": peaks["source"] + ", sustained cuBLAS bf16",
                    "families": {k: {a: round(b, 4) for a, b in v.items()} for k, v in fam.items()},
                    "vit_backbone": {"tflops": round(vit_tflops, 1), "frac": round(vit_tflops / peaks["tflops"], 4),
                                     "ms_per_step": round(vit_ms / args.steps, 3)},
                    "other_ms_per_step": {k: round(prof[k][0] / args.steps, 3)
                                          for k in ("misc", "layernorm", "gemm_other", "head", "smplx")}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(images=1)
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp16 tensor-core operands, fp32 accumulate/residual",
        "data": "synthetic",
        "config": {"workload": f"{w['name']} batch {B}/GPU, synthetic {S}x{S}, random-init ViT-L weights",
                   "images_per_gpu": B, "global_batch": world * B, "persons_in_batch": int(P_last),
                   "det_thresh": w["det_thresh"], "nms_kernel_size": w["nms_kernel_size"],
                   "parallelism": f"dp{world} (image shards, 1 all-gather of person records)" if world > 1 else "dp1",
                   "l2": "working set per step (0.6 GB fp16 weights + >1 GB activations) exceeds the 126 MB L2"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": round(e2e_value, 3), "unit": "images/s", "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": round(ms_e2e / args.steps, 3),
                "persons": int(P_e2e)},
        "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port timed on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_setup(images: int):
    import torch

    from multihmr_b200 import synth
    from oracle import multihmr_ref, smplx_ref

    torch.set_num_threads(os.cpu_count())
    w = WORKLOAD
    sd, bm = build_workload(det_bias=-4.0)
    cfg = multihmr_ref.RefConfig(backbone=w["backbone"], img_size=w["img_size"])
    body = smplx_ref.SMPLXShim(bm, 10)
    x = synth.make_images(images, w["img_size"], seed=w["seed"])
    K = synth.make_cameras(images, w["img_size"], seed=w["seed"])
    idx = synth.make_forced_idx(images, w["img_size"] // 14, w["target_persons_per_image"], seed=w["seed"])

    def forward():
        with torch.no_grad():
            return multihmr_ref.model_forward(sd, body, cfg, x, K, idx=idx, is_training=True)

    return forward


def cpu_baseline(images: int = 1) -> dict:
    fwd = _cpu_setup(images)
    t0 = time.perf_counter()
    fwd()
    dt = time.perf_counter() - t0
    return {"value": round(images / dt, 5), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{images} image of {WORKLOAD['name']} (CPU images/s is batch-independent), "
                      f"{WORKLOAD['target_persons_per_image']} persons/image, fp32 PyTorch oracle port, "
                      f"1 timed forward, {dt:.1f} s"}


def run_reference(args):
    """Reference arm: the reference's own algorithm on the host CPU (the Python reference cannot travel to
    the GPU box, so this is the oracle port pinned against it by oracle/make_golden.py)."""
    world, rank, _ = dist_env()
    if rank != 0:
        return
    fwd = _cpu_setup(1)
    budget_s = 240.0
    t0 = time.perf_counter()
    fwd()  # warm-up (1 step, bounded)
    t_warm = time.perf_counter() - t0
    steps = max(1, min(args.steps, int((budget_s - t_warm) // max(t_warm, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        fwd()
    dt = time.perf_counter() - t0
    value = steps / dt
    sample = (f"each step = 1 image of {WORKLOAD['name']} (of the bs-8 workload; CPU images/s is batch-independent); "
              f"steps capped to {steps} and warm-up to 1 to stay within ~{budget_s:.0f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": round(dt / steps * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD['name']} batch {WORKLOAD['batch_per_gpu']}/GPU, synthetic 896x896, "
                               "random-init ViT-L weights (CPU: 1-image sample per step)"},
        "cpu_baseline": {"value": round(value, 5), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(value, 5), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the (slow) CPU oracle leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
