#!/usr/bin/env python
"""Benchmark of the blurred-GAN hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Workload (BASELINE.json configs[1]): GaussianBlur forward + backward on a synthetic CelebA-128 batch,
256 x 3 x 128 x 128 fp32 NCHW, torch tap rule, std swept over {0.5, 1, 2, 3, 5, 8, 12, 16, 23}.
One *step* = one forward blur (of x) plus one backward blur (of grad_y -- the operator is self-adjoint,
so backward is the same kernel) at every std of the sweep: 18 kernel launches, 9 x 256 images.

  value  images/s, whole job, inputs resident in HBM, CUDA-graph replay of the step's launches through the C ABI,
         timed with CUDA events; buffer sets are rotated so that consecutive steps never re-touch L2-resident data.
  e2e    the same metric through the public module (GaussianBlur(std)(x); y.backward(g)) with HOST buffers:
         pinned host -> device copies of x and grad_y and device -> host copies of y and grad_x inside the timed region.
  roofline      achieved = algorithmic bytes of one launch (read N*C*H*W*4, write N*C*H*W*4) / mean launch time.
  cpu_baseline  the oracle port of the reference's own CPU path (two F.conv2d calls + ATen autograd), all host threads,
                on a bounded sample of the same workload.
  swd    secondary: sliced_wasserstein.API evaluation time on synthetic 128x128x3 images (BASELINE configs[4] shape).

`--impl reference` times only the CPU port and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STD_SWEEP = [0.5, 1.0, 2.0, 3.0, 5.0, 8.0, 12.0, 16.0, 23.0]
BATCH, CH, H, W = 256, 3, 128, 128
WORKLOAD = "cfg2: GaussianBlur fwd+bwd, 256x3x128x128 fp32 NCHW, torch tap rule, std sweep 0.5-23 (9 values)"
METRIC = "GaussianBlur fwd+bwd images/s at 128x128x3"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--no-swd", action="store_true", help="skip the secondary SWD timing")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-large", action="store_true", help="skip the 1024 x 1024 context rows (BASELINE configs[3])")
    p.add_argument("--no-train-step", action="store_true", help="skip the WGAN-GP train-step context row (BASELINE configs[2])")
    p.add_argument("--swd-images", type=int, default=16384, help="images per set for the secondary SWD timing (BASELINE configs[4]: 16384)")
    p.add_argument("--swd-cpu-images", type=int, default=512, help="images per set for the SWD CPU baseline (bounded sample)")
    p.add_argument("--stds", type=str, default="", help="comma-separated override of the std sweep (profiling runs)")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 100 ms while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in {**self.BAD, **self.NOTE}.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.001)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU port of the reference (baseline + reference arm)
# ------------------------------------------------------------------------------------------------
def cpu_fwd_bwd(x, g, std):
    """Reference path on the host: forward = two F.conv2d, backward = ATen autograd of them
    (reference pytorch/gaussianblur.py:63-70; oracle/blur_oracle.py::blur_conv2d_port)."""
    from oracle.blur_oracle import blur_conv2d_port

    xr = x.detach().requires_grad_(True)
    y = blur_conv2d_port(xr, std)
    y.backward(g)
    return y, xr.grad


def cpu_sweep_time(n_images, stds, repeats=1):
    import torch

    torch.manual_seed(0)
    x = torch.rand(n_images, CH, H, W) * 2 - 1
    torch.manual_seed(1)
    g = torch.rand(n_images, CH, H, W) * 2 - 1
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        for std in stds:
            cpu_fwd_bwd(x, g, std)
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_baseline(stds, budget_s=20.0):
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    probe_n = 8
    cpu_sweep_time(probe_n, stds[:1])  # thread-pool / oneDNN warm-up
    t_probe = cpu_sweep_time(probe_n, stds)
    n = int(max(8, min(BATCH, probe_n * budget_s / max(t_probe, 1e-6))))
    n -= n % 8
    t = cpu_sweep_time(n, stds)
    return {"value": n * len(stds) / t, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "%d images x %d stds of the sweep, fwd+bwd, torch %s CPU, %d threads, %.1f s" % (n, len(stds), torch.__version__, cores, t)}


def run_reference_arm(args, stds):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    probe_n = 8
    cpu_sweep_time(probe_n, stds[:1])
    t_probe = cpu_sweep_time(probe_n, stds)
    total_steps = args.steps + args.warmup
    budget_per_step = 150.0 / max(total_steps, 1)
    n = int(max(8, min(BATCH, probe_n * budget_per_step / max(t_probe, 1e-6))))
    n -= n % 8
    for _ in range(args.warmup):
        cpu_sweep_time(n, stds)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sweep_time(n, stds)
    dt = time.perf_counter() - t0
    value = args.steps * n * len(stds) / dt
    sample = "%d images x %d stds per step (bounded sample of the 256-image batch), fwd+bwd, torch %s CPU, %d threads" % (
        n, len(stds), torch.__version__, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "stds": stds, "sample_images_per_step": n, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_swd:
        line["swd"] = {"cpu_baseline": swd_cpu_baseline(args.swd_cpu_images)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def run_native(args, stds):
    import torch
    import torch.distributed as dist

    from blurred_gan_b200 import _lib, ops
    from blurred_gan_b200.pytorch.gaussianblur import GaussianBlur

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    hbm_peak, peak_src = peaks()

    # ---- buffers: enough rotating sets that a step never finds its data in the 126 MB L2 -------------
    n_sets = 5
    plane_bytes = BATCH * CH * H * W * 4
    torch.manual_seed(1000 + rank)
    sets = []
    for _ in range(n_sets):
        x = torch.rand(BATCH, CH, H, W, device=dev) * 2 - 1
        gy = torch.rand(BATCH, CH, H, W, device=dev) * 2 - 1
        sets.append((x, torch.empty_like(x), gy, torch.empty_like(x)))
    plans = [_lib.blur_plan(_lib.RULE_TORCH, s, H, W) for s in stds]  # (k, radius, std_eff)
    stream = torch.cuda.Stream(device=dev)

    def launch_pair(bufs, plan):
        x, y, gy, gx = bufs
        _, radius, std_eff = plan
        sp = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.bgan_blur2d_f32(x.data_ptr(), y.data_ptr(), BATCH, CH, H, W, 0, _lib.RULE_TORCH, std_eff, radius, None, None, 0, sp), "fwd")
        _lib.check(lib.bgan_blur2d_f32(gy.data_ptr(), gx.data_ptr(), BATCH, CH, H, W, 0, _lib.RULE_TORCH, std_eff, radius, None, None, 0, sp), "bwd")

    # CUDA graphs: one per buffer set holding the whole step (18 launches, programmatic-dependent-launch edges between
    # them), replayed from the host with no Python in between; plus one per (set, std) pair for the per-std breakdown.
    paths = []
    with torch.cuda.stream(stream):
        for plan in plans:
            launch_pair(sets[0], plan)  # warm the attribute / occupancy caches outside capture
            paths.append(lib.bgan_blur2d_last_path())  # 4 = tensor-core kernel, 5 = direct FFMA kernel
    stream.synchronize()
    step_graphs, pair_graphs = [], []
    for bufs in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for plan in plans:
                launch_pair(bufs, plan)
        step_graphs.append(g)
        row = []
        for plan in plans:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                launch_pair(bufs, plan)
            row.append(g)
        pair_graphs.append(row)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up, then K timed steps ----------------------------------------------------------------
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step_graphs[i % n_sets].replay()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, min(args.steps, 10))
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in stds] for _ in range(reps)]
    with ClockSampler(local_rank) as clocks:
        with torch.cuda.stream(stream):
            e0.record(stream)
            for i in range(args.steps):
                step_graphs[(args.warmup + i) % n_sets].replay()
            e1.record(stream)
        barrier()
        # per-std breakdown (not part of `value`): same launches, one graph per (set, std), events around each
        with torch.cuda.stream(stream):
            for i in range(reps):
                for j, g in enumerate(pair_graphs[i % n_sets]):
                    ev[i][j][0].record(stream)
                    g.replay()
                    ev[i][j][1].record(stream)
        barrier()
    elapsed_ms = e0.elapsed_time(e1)
    per_rank_ms = [elapsed_ms / args.steps]
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [float(v.item()) / args.steps for v in every]  # names the rank that decides the max
        elapsed_ms = max(float(v.item()) for v in every)
    images_per_step = BATCH * len(stds)
    value = world * args.steps * images_per_step / (elapsed_ms * 1e-3)
    launches = args.steps * len(stds) * 2

    KERNELS = {4: "blur_tc128_kernel (tcgen05 Toeplitz GEMMs, TMEM accumulators, TMA staged)",
               5: "blur_direct128_kernel (FFMA, TMA staged, register-resident separable passes; radius <= 5)"}
    per_std = []
    for j, s in enumerate(stds):
        ms = statistics.mean(ev[i][j][0].elapsed_time(ev[i][j][1]) for i in range(reps))
        gbs = 2 * 2 * plane_bytes / (ms * 1e-3) / 1e9
        taps = 2 * plans[j][1] + 1
        per_std.append({"std": s, "taps": taps, "kernel": KERNELS.get(paths[j], "path %d" % paths[j]).split(" ")[0], "ms_fwd_bwd": round(ms, 4),
                        "images_per_s": round(BATCH / (ms * 1e-3), 1), "hbm_gbs": round(gbs, 1), "hbm_frac": round(gbs / hbm_peak, 4)})
    mean_launch_s = (elapsed_ms * 1e-3) / (args.steps * 2 * len(stds))  # mean over every launch of the timed region
    # Two kernels share the timed region (the sweep crosses the radius where the tensor-core formulation takes over).
    # A graph replay cannot be timed per kernel, so each kernel's launch time is its mean in the per-std breakdown (CUDA
    # events around one forward + backward pair), scaled by (timed-region mean) / (breakdown mean) over all launches.
    breakdown_mean_s = statistics.mean(p["ms_fwd_bwd"] for p in per_std) * 1e-3 / 2
    scale = mean_launch_s / breakdown_mean_s
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "blur_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    by_kernel = []
    for path in sorted(set(paths)):
        rows = [p for p, q in zip(per_std, paths) if q == path]
        t = statistics.mean(p["ms_fwd_bwd"] for p in rows) * 1e-3 / 2 * scale
        ach = 2 * plane_bytes / t / 1e9
        tr = (traffic or {}).get("dram_bytes_per_launch_direct" if path == 5 else "dram_bytes_per_launch")
        by_kernel.append({"bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(ach / hbm_peak, 4),
                          "traffic": tr, "kernel": KERNELS.get(path, "path %d" % path), "launch_us": round(t * 1e6, 2),
                          "launches_per_step": 2 * len(rows), "share_of_step": round(len(rows) * t / (len(per_std) * mean_launch_s), 3),
                          "stds": [p["std"] for p in rows], "algorithmic_bytes_per_launch": 2 * plane_bytes, "peak_source": peak_src})
    dominant = max(by_kernel, key=lambda k: k["share_of_step"])
    roofline = dict(dominant)
    roofline["note"] = ("dominant kernel of the step by time; launch time = its mean in the per-std breakdown scaled to the timed region "
                        "(see bench.py); all launches of the step together: %.1f GB/s = %.4f of peak" % (
                            2 * plane_bytes / mean_launch_s / 1e9, 2 * plane_bytes / mean_launch_s / 1e9 / hbm_peak))
    roofline["all_launches"] = {"achieved": round(2 * plane_bytes / mean_launch_s / 1e9, 1), "frac": round(2 * plane_bytes / mean_launch_s / 1e9 / hbm_peak, 4),
                                "launch_us": round(mean_launch_s * 1e6, 2)}
    roofline["by_kernel"] = by_kernel

    # ---- context: a device-to-device copy of one launch's bytes (what "HBM-bound" means at this size), and the NHWC layout
    copy_ref = None
    nhwc = None
    if world == 1 or rank == 0:
        with torch.cuda.stream(stream):
            for i in range(3):
                sets[i % n_sets][1].copy_(sets[i % n_sets][0])
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            for i in range(20):
                sets[i % n_sets][1].copy_(sets[i % n_sets][0])
            c1.record(stream)
        stream.synchronize()
        cus = c0.elapsed_time(c1) / 20 * 1e3
        copy_ref = {"what": "torch copy_ of one launch's bytes (50.3 MB -> 50.3 MB), rotating buffers", "launch_us": round(cus, 2),
                    "gbs": round(2 * plane_bytes / (cus * 1e-6) / 1e9, 1), "frac": round(2 * plane_bytes / (cus * 1e-6) / 1e9 / hbm_peak, 4)}
        nhwc = []
        xs = [t[0].view(BATCH, H, W, CH) for t in sets]
        ys = [t[1].view(BATCH, H, W, CH) for t in sets]
        for s in (0.5, 1.0, 5.0, 23.0):
            _, radius, std_eff = _lib.blur_plan(_lib.RULE_TF, s, H, W)
            with torch.cuda.stream(stream):
                sp = torch.cuda.current_stream().cuda_stream

                def go(i):
                    _lib.check(lib.bgan_blur2d_f32(xs[i % n_sets].data_ptr(), ys[i % n_sets].data_ptr(), BATCH, CH, H, W, 1, _lib.RULE_TF, std_eff,
                                                   radius, None, None, 0, sp), "nhwc")

                go(0)
                path = lib.bgan_blur2d_last_path()
                c0.record(stream)
                for i in range(20):
                    go(i + 1)
                c1.record(stream)
            stream.synchronize()
            us = c0.elapsed_time(c1) / 20 * 1e3
            nhwc.append({"std": s, "taps": 2 * radius + 1, "kernel": KERNELS.get(path, "path %d" % path).split(" ")[0], "launch_us": round(us, 2),
                         "hbm_frac": round(2 * plane_bytes / (us * 1e-6) / 1e9 / hbm_peak, 4)})

    # ---- context: the critic step's pair of blurs (generated + real batch, same sigma) as ONE launch
    multi = None
    if world == 1 or rank == 0:
        import ctypes

        multi = []
        for s in (5.0, 23.0):
            _, radius, std_eff = _lib.blur_plan(_lib.RULE_TORCH, s, H, W)
            pairs = [((ctypes.c_void_p * 2)(sets[(2 * i) % n_sets][0].data_ptr(), sets[(2 * i + 1) % n_sets][0].data_ptr()),
                      (ctypes.c_void_p * 2)(sets[(2 * i) % n_sets][1].data_ptr(), sets[(2 * i + 1) % n_sets][1].data_ptr())) for i in range(n_sets)]
            with torch.cuda.stream(stream):
                sp = torch.cuda.current_stream().cuda_stream
                res = {}
                for name in ("two_launches", "one_launch"):
                    def go(i):
                        xp, yp = pairs[i % n_sets]
                        if name == "one_launch":
                            _lib.check(lib.bgan_blur2d_multi_f32(xp, yp, 2, BATCH, CH, H, W, 0, _lib.RULE_TORCH, std_eff, radius, None, None, 0, sp), "multi")
                        else:
                            for t in range(2):
                                _lib.check(lib.bgan_blur2d_f32(xp[t], yp[t], BATCH, CH, H, W, 0, _lib.RULE_TORCH, std_eff, radius, None, None, 0, sp), "blur")

                    go(0)
                    c0.record(stream)
                    for i in range(20):
                        go(i + 1)
                    c1.record(stream)
                    stream.synchronize()
                    res[name] = c0.elapsed_time(c1) / 20 * 1e3
            multi.append({"std": s, "tensors": 2, "two_launches_us": round(res["two_launches"], 2), "one_launch_us": round(res["one_launch"], 2),
                          "one_launch_hbm_frac": round(4 * plane_bytes / (res["one_launch"] * 1e-6) / 1e9 / hbm_peak, 4)})

    # ---- context: BASELINE configs[3], the large-image stress [32, 3, 1024, 1024], sigma 8 / 32 / 64 (TF rule: up to 385 taps)
    large = None
    if (world == 1 or rank == 0) and not args.no_large:
        large = []
        LN, LH = 32, 1024
        lbytes = LN * CH * LH * LH * 4
        xl = [torch.rand(LN, CH, LH, LH, device=dev) * 2 - 1 for _ in range(2)]
        yl = [torch.empty_like(xl[0]) for _ in range(2)]
        for s in (8.0, 32.0, 64.0):
            _, radius, std_eff = _lib.blur_plan(_lib.RULE_TF, s, LH, LH)
            ws_bytes = lib.bgan_blur2d_workspace_bytes(LN, CH, LH, LH, 0, radius)
            ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
            with torch.cuda.stream(stream):
                sp = torch.cuda.current_stream().cuda_stream

                def go(i):
                    _lib.check(lib.bgan_blur2d_f32(xl[i & 1].data_ptr(), yl[i & 1].data_ptr(), LN, CH, LH, LH, 0, _lib.RULE_TF, std_eff, radius, None,
                                                   ws.data_ptr(), ws_bytes, sp), "cfg4")

                go(0)
                path = lib.bgan_blur2d_last_path()
                go(1)
                c0.record(stream)
                for i in range(6):
                    go(i)
                c1.record(stream)
            stream.synchronize()
            us = c0.elapsed_time(c1) / 6 * 1e3
            large.append({"std": s, "taps": 2 * radius + 1, "path": path, "launch_ms": round(us * 1e-3, 4),
                          "images_per_s": round(LN / (us * 1e-6), 1), "algorithmic_bytes": 2 * lbytes,
                          "hbm_frac": round(2 * lbytes / (us * 1e-6) / 1e9 / hbm_peak, 4)})
        del xl, yl, ws

    # ---- e2e through the public module with host buffers ----------------------------------------------
    e2e = None
    if True:
        xh = torch.empty(BATCH, CH, H, W).uniform_(-1, 1).pin_memory()
        gh = torch.empty(BATCH, CH, H, W).uniform_(-1, 1).pin_memory()
        yh = torch.empty(BATCH, CH, H, W).pin_memory()
        gxh = torch.empty(BATCH, CH, H, W).pin_memory()
        mod = GaussianBlur(stds[0]).to(dev)

        # Host buffers in, host buffers out.  The blur itself is the public module call; the copies around it are what a
        # caller with host data does, written as a 3-stream pipeline (H2D | blur | D2H, two device buffer sets) so that the
        # two PCIe directions overlap.
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        dbuf = [[torch.empty(BATCH, CH, H, W, device=dev) for _ in range(2)] for _ in range(2)]  # [slot][x | g]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        for e in ev_free:
            e.record(cur)

        def e2e_step():
            for i, s in enumerate(stds):
                slot = i & 1
                mod.set_std(s)
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_free[slot])       # the slot's previous results have been copied out
                    dbuf[slot][0].copy_(xh, non_blocking=True)
                    dbuf[slot][1].copy_(gh, non_blocking=True)
                    ev_in[slot].record(s_in)
                cur.wait_event(ev_in[slot])
                x = dbuf[slot][0].requires_grad_(True)
                y = mod(x)
                y.backward(dbuf[slot][1])
                gx = x.grad
                x.grad = None
                x.requires_grad_(False)
                ev_done[slot].record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[slot])
                    yh.copy_(y.detach(), non_blocking=True)
                    gxh.copy_(gx, non_blocking=True)
                    ev_free[slot].record(s_out)
                y.record_stream(s_out)
                gx.record_stream(s_out)

        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(2):
            e2e_step()
        # seven windows of e2e_steps steps, the median window is reported: the pipeline is driven by the host (copies, module
        # calls, the autograd thread), and a hiccup of a shared host otherwise decides the number (one round-2 run on a busy
        # box had two of three windows 25 % slower than the third)
        windows, windows_local = [], []
        for _ in range(7):
            barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            torch.cuda.synchronize(dev)  # (all three streams)
            dt = time.perf_counter() - t0
            windows_local.append(dt)
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            windows.append(dt)
        dt = sorted(windows)[len(windows) // 2]
        # The ceiling of this pipeline: the same pinned buffers copied host -> device and device -> host on the two copy
        # streams, both directions at once, no kernels -- measured at the same world size in the same run (all ranks copy
        # at the same time, like in the timed windows above).
        def copy_only():
            for i in range(len(stds)):
                slot = i & 1
                with torch.cuda.stream(s_in):
                    dbuf[slot][0].copy_(xh, non_blocking=True)
                    dbuf[slot][1].copy_(gh, non_blocking=True)
                with torch.cuda.stream(s_out):
                    yh.copy_(dbuf[1 - slot][0], non_blocking=True)
                    gxh.copy_(dbuf[1 - slot][1], non_blocking=True)

        copy_only()
        cwin = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                copy_only()
            torch.cuda.synchronize(dev)
            cwin.append(time.perf_counter() - t0)
        cdt_local = sorted(cwin)[1]
        dt_local = sorted(windows_local)[len(windows_local) // 2]
        per_dir = 2 * plane_bytes * len(stds) * e2e_steps  # bytes per direction per rank per window
        stats = torch.tensor([dt_local, cdt_local], device=dev, dtype=torch.float64)
        if world > 1:
            gathered = [torch.zeros_like(stats) for _ in range(world)]
            dist.all_gather(gathered, stats)
        else:
            gathered = [stats]
        per_rank = [g.tolist() for g in gathered]
        cdt = max(r[1] for r in per_rank)
        e2e = {"value": world * e2e_steps * images_per_step / dt, "unit": "images/s",
               "h2d_bytes_per_step": 2 * plane_bytes * len(stds), "d2h_bytes_per_step": 2 * plane_bytes * len(stds),
               "steps": e2e_steps, "windows_s": [round(w, 4) for w in windows],
               "achieved_gbs_per_direction_per_rank": [round(per_dir / r[0] / 1e9, 2) for r in per_rank],
               "copy_ceiling": {"what": "the same pinned buffers, H2D and D2H at once on the two copy streams, no kernels, all ranks at the same time",
                                "gbs_per_direction_per_rank": [round(per_dir / r[1] / 1e9, 2) for r in per_rank],
                                "images_per_s_if_copy_bound": round(world * e2e_steps * images_per_step / cdt, 1)},
               "frac_of_copy_ceiling": round(cdt / dt, 4),
               "api": "GaussianBlur(std)(x); y.backward(g); pinned host x, g -> host y, grad_x; H2D | blur | D2H on three streams"}
        launches += 0  # e2e launches are outside the timed `value` region

    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "ms_per_step_per_rank": [round(v, 5) for v in per_rank_ms],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "stds": stds, "images_per_step": images_per_step, "launches_per_step": 2 * len(stds),
                   "l2": "5 rotating buffer sets (1.0 GB per GPU) > 126 MB L2; one CUDA graph per step",
                   "parallelism": "batch sharded, %d x 256 images, no collective" % world},
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches * world, "roofline": roofline, "per_std": per_std,
        "copy_reference": copy_ref, "nhwc_tf_rule": nhwc, "multi_tensor_launch": multi, "large_image_cfg4": large,
    }

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(stds)
    if not args.no_train_step:  # context: BASELINE configs[2], every rank takes part (gradient all-reduce)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import train_step

            ts = train_step.run(global_batch=256, steps=4, warmup=2, std=5.0, fused_critic=False)
            tf = train_step.run(global_batch=256, steps=4, warmup=2, std=5.0, fused_critic=True)
            line["train_step_cfg3"] = {
                "what": ts["config"]["workload"] + "; the nets are cuDNN fp32 (library code, outside the path); the blur's share is what the path adds",
                "ms_per_step": round(ts["ms_per_step"], 3), "images_per_s": round(ts["value"], 1),
                "ms_per_step_without_blur": round(ts["ms_per_step_without_blur"], 3), "blur_share_of_step": round(ts["blur_share_of_step"], 4),
                "fused_critic_ms_per_step": round(tf["ms_per_step"], 3), "blur_applications_per_step": [7, 6]}
            torch.cuda.empty_cache()
        except Exception as exc:  # context row: never lose the headline line
            line["train_step_cfg3"] = {"error": repr(exc)}
    if not args.no_swd:  # every rank takes part (sharded feed, NCCL exchange)
        try:
            swd = swd_timing(args.swd_images, dev, rank, world, hbm_peak)
        except Exception as exc:  # secondary measurement: never lose the headline line
            swd = {"error": repr(exc)}
        if rank == 0 and not args.no_cpu_baseline and world == 1 and "error" not in swd:
            swd["cpu_baseline"] = swd_cpu_baseline(args.swd_cpu_images)
            swd["speedup_vs_cpu_extrapolated"] = round(swd["cpu_baseline"]["extrapolated_s_at_16384"] * (swd["images_per_set"] / 16384.0) / swd["eval_s"], 1)
        line["swd"] = swd
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


class _EventTimer:
    """Device time of selected calls: a CUDA event pair around every call (no synchronisation; summed after the run)."""

    def __init__(self):
        self.pairs = {}
        self.patched = []

    def wrap(self, mod, name):
        import torch

        fn = getattr(mod, name)
        pairs = self.pairs.setdefault(name, [])

        def timed(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            pairs.append((e0, e1))
            return out

        setattr(mod, name, timed)
        self.patched.append((mod, name, fn))

    def restore(self):
        for mod, name, fn in self.patched:
            setattr(mod, name, fn)
        self.patched = []

    def reset(self):
        for v in self.pairs.values():
            del v[:]

    def seconds(self, name):
        return sum(a.elapsed_time(b) for a, b in self.pairs.get(name, [])) * 1e-3

    def calls(self, name):
        return len(self.pairs.get(name, []))


# HBM bytes a sort moves per key.  Radix passes (csrc/sort.cu): 4 digit passes x (histogram read + scatter read + write) = 48.
# Value buckets (csrc/sort_sample.cu, segments of up to 2^21 keys): classify 4 + 2 (ids), scatter 4 + 2 + 4, local sort 4 + 4
# = 24, plus the per-block bucket tables (6 bytes per (block of 16384 keys, bucket), written and read once).  In the fused
# sort + |A - B| call the last step over B reads the sorted A instead of writing (same count).
SORT_BYTES_PER_KEY = 48


def sort_bytes_per_key(rows_per_segment, path):
    if path != 2:
        return float(SORT_BYTES_PER_KEY)
    buckets = 2
    while buckets * 2048 < rows_per_segment and buckets < 1024:
        buckets *= 2
    return 24.0 + 12.0 * buckets / 16384


def swd_timing(n_images, dev, rank=0, world=1, hbm_peak=None):
    """BASELINE configs[4]: sliced_wasserstein.API on n_images reals + n_images fakes of 3x128x128, the images sharded
    over the ranks (rank r feeds the global minibatches k * world + r), inputs resident in HBM.  With world > 1 the
    channel statistics are all-reduced, the standardised descriptors all-gathered over NCCL and every rank sorts its own
    512 / world directions (blurred_gan_b200/dist.py).  Time = wall clock of the whole evaluation, max over ranks."""
    import torch
    import torch.distributed as dist

    from blurred_gan_b200 import dist as bd
    from blurred_gan_b200 import ops
    from blurred_gan_b200 import sliced_wasserstein as sw

    mb = 64
    rounds = n_images // (mb * world)
    per_rank = rounds * mb
    gen = torch.Generator(device=dev)
    gen.manual_seed(5 + rank)
    reals = torch.rand(per_rank, 3, 128, 128, device=dev, generator=gen) * 2 - 1
    fakes = (torch.rand(per_rank, 3, 128, 128, device=dev, generator=gen) * 2 - 1) * 0.9
    timer = _EventTimer()

    def run(overlap=True):
        np.random.seed(0)
        api = sw.API((3, 128, 128), distributed=world > 1, overlap_gather=overlap)
        t = {}
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        api.begin("reals")
        for i in range(0, per_rank, mb):
            api.feed("reals", reals[i : i + mb])
        api.end("reals")
        api.begin("fakes")
        for i in range(0, per_rank, mb):
            api.feed("fakes", fakes[i : i + mb])
        torch.cuda.synchronize(dev)
        t["feed_s"] = time.perf_counter() - t0
        res = api.end("fakes")
        torch.cuda.synchronize(dev)
        t["total_s"] = time.perf_counter() - t0
        return res, t

    run()  # warm-up (allocator, NCCL channels)
    res, t = run()
    # third run with CUDA-event pairs around the library calls: device time per phase (not part of eval_s)
    for name in ("swd_sort_diff_", "swd_project", "swd_normalise_", "swd_feed"):
        timer.wrap(ops, name)
    timer.wrap(bd, "all_gather_rows")
    try:
        run(overlap=False)  # every all-gather waited for on the spot, so that its event pair brackets it
        torch.cuda.synchronize(dev)
        phases = {name: timer.seconds(name) for name in timer.pairs}
        calls = {name: timer.calls(name) for name in timer.pairs}
    finally:
        timer.restore()
    times = torch.tensor([t["total_s"], t["feed_s"]] + [phases[k] for k in sorted(phases)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    times = times.tolist()
    total_s, feed_s = times[0], times[1]
    phases = dict(zip(sorted(phases), times[2:]))
    rows = rounds * world * mb * 128                   # descriptors per level per set, whole job
    keys_job = 2 * 4 * 4 * 128 * rows                  # A and B, 4 levels, 512 directions (the zero real-vs-real pass is not counted)
    keys_rank = keys_job // world
    from blurred_gan_b200 import _lib
    sort_path = _lib.load().bgan_swd_sort_last_path()
    bpk = sort_bytes_per_key(rows, sort_path)
    desc_bytes_job = 2 * 4 * rows * 147 * 4            # all-gather payload: every rank ends up with all of it
    out = {"images_per_set": rounds * world * mb, "n_gpus": world, "eval_s": round(total_s, 4), "feed_s": round(feed_s, 4),
           "end_s": round(total_s - feed_s, 4), "images_per_s": round(2 * rounds * world * mb / total_s, 1),
           "result_x1e3": [float(v) for v in res],
           "device_time_s": {k: round(v, 4) for k, v in phases.items()}, "calls": calls,
           "sort": {"keys_whole_job": keys_job, "keys_per_rank": keys_rank, "sort_s": round(phases["swd_sort_diff_"], 4),
                    "gkeys_per_s_per_gpu": round(keys_rank / max(phases["swd_sort_diff_"], 1e-9) / 1e9, 2),
                    "gkeys_per_s_whole_job": round(keys_job / max(phases["swd_sort_diff_"], 1e-9) / 1e9, 2),
                    "keys_per_segment": rows,
                    "path": {1: "radix passes", 2: "value buckets", 3: "value buckets gave up, radix passes"}.get(sort_path, str(sort_path)),
                    "hbm_bytes_per_key": round(bpk, 2),
                    "hbm_gbs": round(bpk * keys_rank / max(phases["swd_sort_diff_"], 1e-9) / 1e9, 1),
                    "hbm_frac": round(bpk * keys_rank / max(phases["swd_sort_diff_"], 1e-9) / 1e9 / hbm_peak, 4) if hbm_peak else None},
           "note": "host RNG draws (reference order) are inside eval_s; device_time_s: CUDA-event pairs around the calls in a separate run; "
                   "swd_sort_diff_ = both sorts + the fused |A - B| sums (bgan_swd_sort_diff_f32)"}
    if world > 1:
        ag = phases["all_gather_rows"]
        recv = desc_bytes_job * (world - 1) // world
        out["all_gather"] = {"payload_bytes_whole_set": desc_bytes_job, "received_bytes_per_rank": recv, "device_time_s": round(ag, 4),
                             "gbs_received_per_rank": round(recv / max(ag, 1e-9) / 1e9, 1),
                             "note": "torch.distributed all_gather_into_tensor over NCCL, 8 calls (4 levels x 2 sets); device time incl. the wait for the slowest rank, "
                                     "measured in the separate phase run with the gathers waited for on the spot; in eval_s they run asynchronously beside the "
                                     "next set's feed kernels and the previous level's sort"}
    return out


def swd_cpu_baseline(n_images=512):
    """The reference's own path (oracle restatement of sliced_wasserstein.API, numpy) on the host, on a bounded sample."""
    from oracle import swd_oracle as so

    rs = np.random.RandomState(5)
    reals = (rs.rand(n_images, 3, 128, 128) * 2 - 1).astype(np.float32)
    fakes = ((rs.rand(n_images, 3, 128, 128) * 2 - 1) * 0.9).astype(np.float32)
    np.random.seed(0)
    t0 = time.perf_counter()
    api = so.API((3, 128, 128))
    for mode, data in (("reals", reals), ("fakes", fakes)):
        api.begin(mode)
        for i in range(0, n_images, 64):
            api.feed(mode, data[i : i + 64].copy())
        res = api.end(mode)
    dt = time.perf_counter() - t0
    return {"value": dt, "unit": "s", "images_per_set": n_images, "images_per_s": round(2 * n_images / dt, 2), "cores": os.cpu_count() or 1,
            "kind": "port", "result_x1e3": [float(v) for v in res],
            "sample": "oracle/swd_oracle.API (numpy restatement of sliced_wasserstein.py:98-133) on 2 x %d images; numpy's sort is single-threaded, "
                      "BLAS matmul uses the host threads" % n_images,
            "extrapolated_s_at_16384": round(dt * 16384 / n_images, 1),
            "extrapolation": "linear in the image count (a lower bound: the sort, ~90 %% of it, grows as n log n)"}


def main():
    args = parse_args()
    stds = [float(s) for s in args.stds.split(",")] if args.stds else STD_SWEEP
    if args.impl == "reference":
        run_reference_arm(args, stds)
    else:
        run_native(args, stds)


if __name__ == "__main__":
    main()
