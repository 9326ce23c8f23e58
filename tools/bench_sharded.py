#!/usr/bin/env python
"""BASELINE.json configs 3-5 as sharded batches: every rank owns a contiguous block of the item list
(fb_batch_shard), runs the device-resident batch entry point on its own GPU and the results are gathered in input
order — float64 scores for MS-SSIM, NRGBA outputs for Lanczos / blur+sharpen.  Weak scaling (fixed items per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py [--items K]

Rank 0 prints one JSON line per config: whole-job items/s with and without the gather, max over ranks, CUDA events.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fennec_b200 import api, batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--items", type=int, default=8, help="items per GPU per step")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
api.set_device(local)
n = args.items
total = n * world
lo, hi = batch.shard_range(total, world, rank)
assert hi - lo == n


def noise(k, h, w, seed):
    g = torch.Generator(device="cuda").manual_seed(seed + 1000 * rank)
    t = torch.randint(0, 256, (k, h, w, 4), dtype=torch.uint8, device="cuda", generator=g)
    t[..., 3] = 255
    return t


def timed(fn, steps):
    for _ in range(3):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


_gbuf = {}


def gather_images(x):
    """Input-order gather of equal shards: ncclAllGather straight into one preallocated tensor (the list form of
    dist.all_gather adds a device copy per shard and torch.cat another)."""
    if world == 1:
        return x
    key = (tuple(x.shape), x.dtype)
    if key not in _gbuf:
        _gbuf[key] = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(_gbuf[key], x)
    return _gbuf[key]


def line(name, ms_compute, ms_total, extra):
    if rank == 0:
        print(json.dumps({"config": name, "n_gpus": world, "items_per_gpu": n, "ms_per_step_compute": round(ms_compute, 4),
                          "ms_per_step_with_gather": round(ms_total, 4), "items_per_s_compute": round(total / ms_compute * 1e3, 1),
                          "items_per_s_with_gather": round(total / ms_total * 1e3, 1), **extra}), flush=True)


# config 3: GaussianBlur sigma=2 + Sharpen 0.5 on 3840x2160
x = noise(n, 2160, 3840, 1)
y, z = torch.empty_like(x), torch.empty_like(x)
def c3():
    batch.gaussian_blur_batch(x, 2.0, out=y); batch.sharpen_batch(y, 0.5, out=z)
line("3: blur s=2 + sharpen 0.5, 3840x2160", timed(c3, args.steps), timed(lambda: (c3(), gather_images(z)), args.steps),
     {"gather": "all_gather of NRGBA outputs (33.2 MB per item)"})
del x, y, z
# config 4: Lanczos-3 7680x4320 -> 1920x1080
x = noise(n, 4320, 7680, 2)
out = torch.empty((n, 1080, 1920, 4), dtype=torch.uint8, device="cuda")
c4 = lambda: batch.lanczos_resize_batch(x, 1920, 1080, out=out)  # noqa: E731
line("4: Lanczos-3 7680x4320 -> 1920x1080", timed(c4, args.steps), timed(lambda: (c4(), gather_images(out)), args.steps),
     {"gather": "all_gather of NRGBA outputs (8.3 MB per item)"})
# configs 3 and 4 with the gather FUSED into the producing kernel: the destination of rank r's batch is its slice of a
# gathered buffer that lives on rank 0 (symmetric memory, mapped over NVLink; batch.PeerGather), so the kernel's own
# stores are the transfer — no collective, no second pass over the outputs.
if world > 1:
    try:
        pg = batch.PeerGather(total, (1080, 1920, 4), world, rank)
        mine = pg.my_slice()
        c4_fused = lambda: batch.lanczos_resize_batch(x, 1920, 1080, out=mine)  # noqa: E731
        ms = timed(lambda: (c4_fused(), pg.handle.barrier()), args.steps)
        c4(); ref_all = gather_images(out).clone()
        c4_fused(); got = pg.finish()
        ok = bool(torch.equal(got, ref_all)) if rank == 0 else True
        line("4: Lanczos-3, gather fused into the kernel's stores (peer memory on rank 0)", timed(c4, args.steps), ms,
             {"gather": "P2P stores over NVLink into rank 0's buffer; equals the all_gather result: %s" % ok})
        del pg, mine, ref_all
        x4k = noise(n, 2160, 3840, 1)
        y4k = torch.empty_like(x4k)
        pg3 = batch.PeerGather(total, (2160, 3840, 4), world, rank)
        mine3 = pg3.my_slice()
        def c3_fused():
            batch.gaussian_blur_batch(x4k, 2.0, out=y4k); batch.sharpen_batch(y4k, 0.5, out=mine3)
        ms3 = timed(lambda: (c3_fused(), pg3.handle.barrier()), args.steps)
        line("3: blur + sharpen, gather fused into the sharpen kernel's stores", 0.0 + timed(lambda: (batch.gaussian_blur_batch(x4k, 2.0, out=y4k), batch.sharpen_batch(y4k, 0.5, out=x4k)), args.steps), ms3,
             {"gather": "P2P stores over NVLink into rank 0's buffer (33.2 MB per item)"})
        del pg3, mine3, x4k, y4k
    except Exception as e:  # symmetric memory unavailable
        if rank == 0:
            print(json.dumps({"config": "fused gather", "unavailable": repr(e)[:300]}), flush=True)
del out
# config 5: MS-SSIM on 7680x4320 pairs
b = noise(n, 4320, 7680, 3)
scores = torch.empty(n, dtype=torch.float64, device="cuda")
c5 = lambda: batch.msssim_batch(x, b, out=scores)  # noqa: E731
line("5: MS-SSIM 7680x4320 pairs", timed(c5, args.steps), timed(lambda: (c5(), batch.gather_scores(scores, total, world, rank)), args.steps),
     {"gather": "all_gather of float64 scores"})
if world > 1:
    dist.destroy_process_group()
