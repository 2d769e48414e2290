#!/usr/bin/env python
"""bench.py — Mrays/s of the NeRFshop render path at 1920x1080 on the synthetic nerf/fox-shaped workload.

  python bench.py --gpus N --steps K --warmup W [--impl native|reference]

A "step" is one frame: Testbed::render_nerf of one camera of a 120-view orbit (a different camera every
step = free-viewpoint orbit), base.json network (hash L=16 F=2 T=2^19, MLPs 64x1 / 64x2), synthetic seeded
parameters (no checkpoint ships with the reference). N > 1: image-plane tiles are partitioned across ranks and
one NCCL all-gather assembles the framebuffer inside the step.

value      : whole-job Mrays/s with the output staying in HBM (per-step CUDA-event time, max over ranks)
e2e.value  : the same through the host-buffer entry point nsb_render_host: frame description host->device,
             RGBA+depth framebuffer device->pinned host inside the timed region
roofline   : the one kernel of the step (k_render_fused): algorithmic 512 B hash-grid gather per sample
              x samples of the frame / its CUDA-event duration, against measured HBM peak
cpu_baseline / --impl reference: the CPU oracle (port of the reference path; the reference itself cannot be
             built here) on all host cores on a bounded sample (a 1/64-resolution frame of the same orbit).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
CPU_W, CPU_H = 240, 135  # bounded CPU sample: 1/64 of the pixels of the same camera
N_ORBIT = 120
METRIC = "Mrays/s @1920x1080 nerf/fox render (synthetic params)"
BYTES_PER_SAMPLE = 512   # 16 levels x 8 corners x 2 fp16
FLOP_PER_SAMPLE = 20480  # both MLPs


def load_traffic():
    """DRAM bytes per launch of k_render_fused from the committed `ncu --set full` digest (profiles/), or None."""
    import glob
    import re

    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_render_fused*.txt"))):
        txt = open(path).read()
        rd = re.search(r"dram__bytes_read\.sum\s+([0-9.]+)\s+(\w+)", txt)
        wr = re.search(r"dram__bytes_write\.sum\s+([0-9.]+)\s+(\w+)", txt)
        if rd and wr:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            best = (float(rd.group(1)) * scale.get(rd.group(2), 1.0) + float(wr.group(1)) * scale.get(wr.group(2), 1.0), os.path.basename(path))
    return best


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(steps: int, warmup: int, threads: int | None = None):
    """The CPU arm: oracle port of the reference path, all host cores, bounded sample per step."""
    from nerfshop_b200 import synthetic as syn
    from oracle import oracle as orc

    model = syn.make_model(seed=1337)
    occ = syn.make_occupancy(model)
    o = orc.Oracle(model.desc, model.params, occ)
    cores = orc.set_threads(threads or (os.cpu_count() or 1))
    cams = syn.orbit_cameras(N_ORBIT)
    times, samples = [], 0
    for i in range(warmup + steps):
        f = syn.make_frame(model, cams[(i * 7) % N_ORBIT], CPU_W, CPU_H)
        t0 = time.perf_counter()
        _, _, st, _ = o.render(f)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            samples += st.n_samples
    total = sum(times)
    mrays = CPU_W * CPU_H * len(times) / total / 1e6
    return mrays, total / len(times) * 1e3, cores, samples / max(len(times), 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "nerf/fox 1080p free-viewpoint orbit, hash L=16 F=2 T=2^19, MLP 64x1 + 64x2 (configs[1])", "resolution": [W, H],
              "cameras": f"{N_ORBIT}-view orbit, one camera per step", "parallelism": f"image-tile partition x{world}" if world > 1 else "single GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 8))
        mrays, ms, cores, spf = cpu_reference_run(steps, min(args.warmup, 1))
        sample = f"{CPU_W}x{CPU_H} frame (1/64 of the 1080p pixels) of the same orbit per step, {cores} threads"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": mrays, "unit": "Mrays/s", "n_gpus": 0, "steps": steps, "warmup": min(args.warmup, 1),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16/f32", "data": "synthetic",
            "config": dict(config, note="CPU oracle port of the reference path; the reference cannot be built here (tiny-cuda-nn/Eigen submodules absent)"),
            "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch
    import torch.distributed as dist

    from nerfshop_b200 import abi, parallel, synthetic as syn
    from nerfshop_b200.renderer import NerfRenderer

    assert torch.cuda.is_available(), "bench.py needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    model = syn.make_model(seed=1337)
    occ = syn.make_occupancy(model)
    r = NerfRenderer(local_rank)
    r.upload_model(model.desc, model.params)
    r.upload_occupancy(occ)
    cams = syn.orbit_cameras(N_ORBIT)

    fb = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    depth = torch.zeros((H, W), dtype=torch.float32, device=dev)
    n_tiles = [r.tiles_for_rank(W, H, k, world) for k in range(world)]
    max_tiles = max(n_tiles)
    if world > 1:
        shard = torch.zeros((max_tiles * 128, 4), dtype=torch.float32, device=dev)
        gathered = torch.zeros((world, max_tiles * 128, 4), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    host_fb = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
    host_depth = torch.zeros((H, W), dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream(dev)
    launches_per_step = 2 + (world if world > 1 else 0)  # k_prepare_rays + k_render_fused (+ k_pack_tiles and world-1 x unpack)

    def device_step(i):
        """One frame, output left in HBM (for N > 1: render own tiles, pack, all-gather, unpack every shard)."""
        f = syn.make_frame(model, cams[(i * 7) % N_ORBIT], W, H, rank=rank, world=world)
        fb.zero_()  # render_buffer.clear_frame
        r.render(f, fb, depth)
        if world > 1:
            parallel.gather_framebuffer(r, fb, rank, world, shard, gathered)

    def timed(step_fn, n_warm, n_steps):
        for i in range(n_warm):
            step_fn(i)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        evs = []
        kern_ms, samples = [], []
        wall0 = time.perf_counter()
        for i in range(n_steps):
            flush.zero_()  # L2 flush between timed iterations, outside the per-step event bracket
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fn(n_warm + i)
            e1.record(stream)
            evs.append((e0, e1))
            st = r.stats()  # synchronises on the render kernel's own events
            kern_ms.append(st.fused_ms)
            samples.append(st.n_samples)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - wall0
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, kern_ms, samples

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, wall, kern_ms, samples = timed(device_step, args.warmup, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers through nsb_render_host (single GPU) / gather then D2H on rank 0 (N > 1) ----
    def e2e_step(i):
        f = syn.make_frame(model, cams[(i * 7) % N_ORBIT], W, H, rank=rank, world=world)
        if world == 1:
            r.render_to_cpu(f, host_fb, host_depth)
        else:
            device_step(i)
            if rank == 0:
                host_fb.copy_(fb, non_blocking=True)
                host_depth.copy_(depth, non_blocking=True)
            torch.cuda.synchronize(dev)

    def timed_wall(step_fn, n_warm, n_steps):
        for i in range(n_warm):
            step_fn(i)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(n_steps):
            step_fn(n_warm + i)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e2e_s = timed_wall(e2e_step, max(3, args.warmup // 2), args.steps)

    if rank == 0:
        hbm_gbs, tflops, peak_kind = load_peaks()
        traffic = load_traffic()
        ms_per_step = total_ms / args.steps
        value = W * H / (ms_per_step * 1e-3) / 1e6
        e2e_value = W * H * args.steps / e2e_s / 1e6
        k_ms = float(np.mean(kern_ms))
        s_mean = float(np.mean(samples))  # this rank's samples per frame
        achieved = s_mean * BYTES_PER_SAMPLE / (k_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16 (fp32 accumulate/composite)",
            "data": "synthetic", "fps": 1e3 / ms_per_step,
            "config": dict(config, l2="flushed between timed steps (256 MiB memset outside the per-step event bracket)",
                           samples_per_frame=s_mean * world, samples_per_ray=s_mean * world / (W * H)),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(np.dtype(np.uint8).itemsize * __import__("ctypes").sizeof(abi.NsbFrame)),
                    "d2h_bytes_per_step": W * H * 20, "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"kernel": "k_render_fused", "bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                         "traffic": traffic[0] if traffic else None, "traffic_source": traffic[1] if traffic else None, "peak_source": peak_kind, "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step if world == 1 else None, "samples_per_launch": s_mean,
                         "tensor_tflops": s_mean * FLOP_PER_SAMPLE / (k_ms * 1e-3) / 1e12, "tensor_frac": s_mean * FLOP_PER_SAMPLE / (k_ms * 1e-3) / 1e12 / tflops},
        }
        if world == 1 and not args.no_cpu_baseline:
            mrays, ms, cores, _ = cpu_reference_run(2, 0)
            out["cpu_baseline"] = {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                   "sample": f"2 frames of {CPU_W}x{CPU_H} (1/64 of the 1080p pixels) of the same orbit, {cores} threads"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    r.close()


if __name__ == "__main__":
    main()
