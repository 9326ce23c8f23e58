#!/usr/bin/env python
"""bench.py — the headline metric of BASELINE.json: 4K fennec.SSIM megapixels/s on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (fennec.SSIM, ssim.go:24-43) over one batch of P synthetic
3840x2160 NRGBA pairs per GPU (default P = 64).  Megapixels count the pixels of ONE image of each pair.

  value     whole-job MP/s with the inputs already resident in HBM (device-resident C-ABI entry point,
            fb_ssim_batch_dev), CUDA-event timed on the launching stream, max over ranks.
  e2e       the same metric through the reference-facing host-buffer call (fb_ssim, what the cgo shim
            binds): pinned HOST buffers, H2D of both images and D2H of the score inside the timed region.
  roofline  achieved algorithmic HBM GB/s of the dominant kernel (8 B per pixel) vs the measured copy peak.
  cpu_baseline  the CPU oracle (a C restatement of the reference's Go path — no Go toolchain exists in
            this image) timed on the box's host cores on a bounded sample of the same workload.

--impl reference times that CPU restatement alone, on the same config/metric (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 3840, 2160
MP_PER_PAIR = W * H / 1e6
BYTES_PER_PAIR = 2 * W * H * 4  # algorithmic traffic: both NRGBA images read once, 8 B per pixel
METRIC = "4k_ssim_megapixels_per_s"
WORKLOAD = "fennec.SSIM (ssim.go:24) on 3840x2160 synthetic NRGBA pairs"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="4K pairs per GPU per step (64 pairs = 4.2 GB >> L2; the grid tail costs 3.4 %% at 32 pairs, 1.7 %% at 64)")
    ap.add_argument("--e2e-pairs", type=int, default=8, help="pairs per e2e step (host buffers)")
    ap.add_argument("--cpu-pairs", type=int, default=0, help="pairs in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ssim_kernel_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        busy = [x for x in sm if x > 500] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_CPU_IMGS = []


def cpu_reference_run(pairs: int, threads: int):
    """Time the CPU restatement of fennec.SSIM (oracle/fennec_oracle.c, the reference's goroutine row
    split restated with pthreads) on `pairs` 4K pair evaluations, cycling over 4 distinct pairs.
    Returns (MP/s, seconds, scores)."""
    from fennec_b200 import synth
    from oracle import pyoracle as O
    O.set_procs(threads)
    while len(_CPU_IMGS) < 4:
        i = len(_CPU_IMGS)
        a = synth.noise_image(W, H, 1000 + i)
        _CPU_IMGS.append((a, synth.perturb(a, 2000 + i, 6)))
    t0 = time.perf_counter()
    scores = [O.ssim(*_CPU_IMGS[i % 4]) for i in range(pairs)]
    dt = time.perf_counter() - t0
    return pairs * MP_PER_PAIR / dt, dt, scores


def run_reference(args):
    """--impl reference: the reference's CPU path (C restatement: kind 'port') on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    threads = os.cpu_count() or 1
    pairs = args.cpu_pairs or 16   # ~2.5 s per step on 16 cores: a bounded sample of the 32-pair GPU step
    cpu_reference_run(2, threads)  # warm the page cache / threads
    times = []
    for _ in range(max(1, min(args.steps, 6))):
        mps, dt, _ = cpu_reference_run(pairs, threads)
        times.append(dt)
    dt = float(np.mean(times))
    value = pairs * MP_PER_PAIR / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "MP/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": pairs, "w": W, "h": H,
                   "note": "CPU restatement of the reference Go path (oracle/fennec_oracle.c); Go toolchain unavailable"},
        "cpu_baseline": {"value": value, "unit": "MP/s", "cores": threads, "kind": "port",
                         "sample": f"{pairs} pairs of 3840x2160 per step, {len(times)} steps"},
        "e2e": {"value": value, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def make_device_batch(torch, pairs: int, seed: int):
    """Seeded uniform-noise pairs generated on the device: b = clip(a + U{-6..6}) on RGB, alpha 255."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randint(0, 256, (pairs, H, W, 4), dtype=torch.uint8, device="cuda", generator=g)
    a[..., 3] = 255
    b = a.clone()
    d = torch.randint(-6, 7, (pairs, H, W, 3), dtype=torch.int16, device="cuda", generator=g)
    b[..., :3] = torch.clamp(d.add_(a[..., :3]), 0, 255).to(torch.uint8)
    del d
    return a, b


def bind_to_gpu_numa_node(index: int):
    """Pin this process (and the pinned host buffers it allocates afterwards) to the CPUs NVML reports as local to
    GPU `index`.  With 8 ranks the e2e leg is host-memory bound: buffers that land on the far socket halve the H2D
    rate.  Returns the number of CPUs bound, or None if NVML has no answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from fennec_b200 import api, batch, build

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    build.build()
    api.set_device(local)

    P = args.pairs
    a, b = make_device_batch(torch, P, 1234 + rank)
    scores = torch.empty(P, dtype=torch.float64, device="cuda")
    total_items = P * world

    def step():
        batch.ssim_batch(a, b, out=scores)
        if world > 1:  # the path's only exchange: gather the per-shard scores in input order
            return batch.gather_scores(scores, total_items, world, rank)
        return scores

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    batch.take_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # per-launch duration of the dominant kernel: events bracketing each fb_ssim_batch_dev enqueue
    ks = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    e0.record()
    for i in range(args.steps):
        ks[i][0].record()
        batch.ssim_batch(a, b, out=scores)
        ks[i][1].record()
        if world > 1:
            batch.gather_scores(scores, total_items, world, rank)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = batch.take_launch_count()
    kernel_ms = float(np.mean([s.elapsed_time(e) for s, e in ks]))
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = total_items * MP_PER_PAIR / (ms_step * 1e-3)

    # ---- e2e: host buffers through the reference-facing call (fb_ssim), copies inside the timed region ----
    E = args.e2e_pairs
    host_a = torch.empty((E, H, W, 4), dtype=torch.uint8).pin_memory()
    host_b = torch.empty((E, H, W, 4), dtype=torch.uint8).pin_memory()
    host_a.copy_(a[:E].cpu())
    host_b.copy_(b[:E].cpu())
    na, nb = host_a.numpy(), host_b.numpy()
    pool = ThreadPoolExecutor(max_workers=4)  # CompressBatch-style concurrent callers; one stream per thread

    def e2e_worker(i):
        api.set_device(local)
        return api.SSIM(na[i], nb[i])

    def e2e_step():
        return list(pool.map(e2e_worker, range(E)))

    e2e_scores = e2e_step()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * E * e2e_steps * MP_PER_PAIR / float(te.item())
    dev_scores = scores[:E].cpu().numpy()
    assert np.all(np.abs(np.array(e2e_scores) - dev_scores) <= 2e-7), "host-buffer and device-resident paths disagree"

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = P * BYTES_PER_PAIR / (kernel_ms * 1e-3) / 1e9
        traffic = recorded_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": P, "w": W, "h": H,
                       "l2": f"inputs larger than L2 ({P * BYTES_PER_PAIR / 1e6:.0f} MB per step per GPU)",
                       "parallelism": f"item-sharded x{world}, all_gather of float64 scores" if world > 1 else "single GPU",
                       "timing": "CUDA events on the launching stream, max over ranks"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic["dram_bytes_per_launch"] * P / traffic.get("pairs_per_launch", P)) if traffic else None,
                         "peak_source": peak_src, "kernel": "ssim_strip_kernel<4> (+32-thread finalize)", "fma_pipe_note": "FP32-FMA-pipe-bound stencil, not HBM-bound: see profiles/ and DESIGN.md K1",
                         "kernel_ms_per_launch": kernel_ms, "algorithmic_bytes_per_launch": P * BYTES_PER_PAIR,
                         },
            "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": E * BYTES_PER_PAIR,
                    "d2h_bytes_per_step": E * 8, "api": "fb_ssim (host buffers, pinned), 4 caller threads",
                    "pairs_per_step": E, "steps": e2e_steps,
                    "host_binding": (f"each rank bound to the {numa_cpus} CPUs NVML reports local to its GPU" if numa_cpus else "none")},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cp = args.cpu_pairs or 64   # ~10 s of wall time on 16 cores
            cpu_reference_run(2, threads)
            mps, secs, cscores = cpu_reference_run(cp, threads)
            line["cpu_baseline"] = {"value": mps, "unit": "MP/s", "cores": threads, "kind": "port",
                                    "sample": f"{cp} evaluations of 3840x2160 pairs (4 distinct, {secs:.1f} s wall), oracle/fennec_oracle.c "
                                              f"(C restatement of the Go path; Go toolchain unavailable)"}
        print(json.dumps(line), flush=True)
    pool.shutdown()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
