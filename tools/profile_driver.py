"""Tiny driver for `ncu --set full` captures: runs one op of the hot path a few times on device-resident
synthetic data and nothing else (keeps the replayed launch count small).

    python tools/profile_driver.py ssim|ssim_fast|msssim|blur|sharpen|adaptive|lanczos|box|ycbcr|analyze|orient|palette [--pairs P] [--iters I]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fennec_b200 import batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("op")
ap.add_argument("--pairs", type=int, default=8)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--w", type=int, default=3840)
ap.add_argument("--h", type=int, default=2160)
ap.add_argument("--opaque", action="store_true", help="alpha = 255 everywhere (the photo case)")
args = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randint(0, 256, (args.pairs, args.h, args.w, 4), dtype=torch.uint8, device="cuda", generator=g)
b = torch.randint(0, 256, (args.pairs, args.h, args.w, 4), dtype=torch.uint8, device="cuda", generator=g)
if args.opaque:
    a[..., 3] = 255
    b[..., 3] = 255
torch.cuda.synchronize()
for _ in range(args.iters):
    if args.op == "ssim":
        r = batch.ssim_batch(a, b)
    elif args.op == "ssim_fast":
        r = batch.ssim_fast_batch(a, b)
    elif args.op == "msssim":
        r = batch.msssim_batch(a, b)
    elif args.op == "blur":
        r = batch.gaussian_blur_batch(a, 2.0)
    elif args.op == "sharpen":
        r = batch.sharpen_batch(a, 0.5)
    elif args.op == "adaptive":
        r = batch.adaptive_sharpen_batch(a, 0.5)
    elif args.op == "analyze":
        r = torch.empty(args.pairs * 2048, dtype=torch.uint8, device="cuda")
        batch.analyze_scan_batch(a, r)
    elif args.op == "ycbcr":
        y = a[..., 0].contiguous(); cb = b[:, ::2, ::2, 1].contiguous(); cr = b[:, ::2, ::2, 2].contiguous()
        r = batch.ycbcr_to_nrgba_batch(y, cb, cr, 2)
    elif args.op == "orient":
        r = batch.apply_orientation_batch(a, 6)
        r = batch.apply_orientation_batch(a, 2)
    elif args.op == "palette":
        pal = b[:, 0, :256, :].contiguous(); pal[..., 3] = 255
        r = batch.apply_palette_batch(a, pal, 256)[1]
    elif args.op == "lanczos":
        r = batch.lanczos_resize_batch(a, args.w // 4, args.h // 4)
    elif args.op == "box":
        r = batch.box_downsample_batch(a, 512, 288)
    else:
        raise SystemExit(f"unknown op {args.op}")
torch.cuda.synchronize()
print(args.op, "done", tuple(r.shape))
