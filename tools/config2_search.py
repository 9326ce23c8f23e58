#!/usr/bin/env python
"""BASELINE.json config 2 as a workload: the SSIM-guided JPEG quality search of compressJPEGOptimal (compress.go:21-88)
on one 4032x3024 image at Balanced (target SSIM 0.94), with ONLY the SSIM step on the GPU.

The search itself stays on the Go side of the boundary (north_star); this harness restates its control flow so the GPU
scorer can be exercised inside it: lo/hi bounds and the fast-path start (compress.go:33-44), encode at mid -> decode ->
SSIMFast(src, decoded) -> keep / raise (compress.go:45-74).  The codec here is libjpeg through Pillow, standing in for
Go's image/jpeg (absent without a Go toolchain) — it only produces the candidate images; both scorers see the same ones:

  oracle   SSIMFast of the CPU oracle on convertToNRGBA(decoded planes)   (ssim.go:48-70, convert.go:34-64 restated)
  gpu      the device session fb_ssim_ref_*: src's thumbnail cached on the device, every iteration uploads the decoded
           Y/Cb/Cr planes (fb_ssim_ref_score_ycbcr) — the binding INTEGRATION.md gives for compress.go:59-62

    python tools/config2_search.py [--w 4032 --h 3024 --target 0.94] [--no-oracle]
prints one JSON object: chosen quality and the per-iteration (Q, score) trace of each scorer, agreement, and timings.
"""
from __future__ import annotations

import argparse
import io
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def decode_planes(jpeg_bytes: bytes):
    """jpeg.Decode stand-in: full-resolution Y, Cb, Cr planes (libjpeg upsamples chroma: ratio 4:4:4)."""
    from PIL import Image
    im = Image.open(io.BytesIO(jpeg_bytes))
    im.draft("YCbCr", im.size)
    if im.mode != "YCbCr":
        im = im.convert("YCbCr")
    arr = np.asarray(im)
    return (np.ascontiguousarray(arr[..., 0]), np.ascontiguousarray(arr[..., 1]), np.ascontiguousarray(arr[..., 2]))


def encode_jpeg(src: np.ndarray, quality: int) -> bytes:
    """encodeJPEG stand-in (io.go:157): opaque NRGBA -> RGB JPEG at `quality`, 4:2:0."""
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(src[..., :3]).save(buf, format="JPEG", quality=int(quality), subsampling=2)
    return buf.getvalue()


def quality_search(src: np.ndarray, target: float, score, codec_cache=None):
    """compressJPEGOptimal's loop (compress.go:21-88).  score(y, cb, cr) -> SSIMFast(src, decoded)."""
    if target >= 1.0:
        target = 0.999                      # compress.go:24-26
    lo, hi = 1, 100
    if target >= 0.99:                      # compress.go:36-44
        lo = 75
    elif target >= 0.97:
        lo = 50
    elif target >= 0.94:
        lo = 30
    elif target >= 0.90:
        lo = 15
    best_q, best_ssim, best_len, trace = hi, 1.0, None, []
    while lo <= hi:
        mid = (lo + hi) // 2
        if codec_cache is not None and mid in codec_cache:
            data, planes = codec_cache[mid]
        else:
            data = encode_jpeg(src, mid)
            planes = decode_planes(data)
            if codec_cache is not None:
                codec_cache[mid] = (data, planes)
        s = score(*planes)
        trace.append((mid, s))
        if s >= target:                     # compress.go:64-70
            best_q, best_ssim, best_len = mid, s, len(data)
            hi = mid - 1
        else:
            lo = mid + 1
    return best_q, best_ssim, best_len, trace


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=4032)
    ap.add_argument("--h", type=int, default=3024)
    ap.add_argument("--target", type=float, default=0.94)
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--no-oracle", action="store_true")
    args = ap.parse_args()
    from fennec_b200 import api, synth
    src = synth.gradient_noise_image(args.w, args.h, args.seed)
    cache = {}
    out = {"config": f"CompressBytes {args.w}x{args.h} at target SSIM {args.target} (compress.go:21-88), SSIM step on the GPU",
           "codec": "libjpeg via Pillow (stand-in for Go image/jpeg: candidates only)"}
    with api.SSIMReference(src) as ref:
        t0 = time.perf_counter()
        q, s, n, tr = quality_search(src, args.target, lambda y, cb, cr: ref.score_ycbcr(y, cb, cr, 0), cache)
        t_first = time.perf_counter() - t0
        t0 = time.perf_counter()
        ms = []
        for mid, _ in tr:                   # the scoring step alone, candidates already decoded
            t1 = time.perf_counter()
            ref.score_ycbcr(*cache[mid][1], 0)
            ms.append((time.perf_counter() - t1) * 1e3)
    out["gpu"] = {"quality": q, "ssim": s, "jpeg_bytes": n, "trace": tr, "iterations": len(tr),
                  "search_wall_s_with_codec": round(t_first, 3), "score_ms_per_iteration": round(float(np.median(ms)), 3),
                  "h2d_bytes_per_iteration": int(3 * args.w * args.h)}
    if not args.no_oracle:
        from oracle import pyoracle as O
        t0 = time.perf_counter()
        q2, s2, n2, tr2 = quality_search(src, args.target, lambda y, cb, cr: O.ssim_fast(src, O.ycbcr_to_nrgba(y, cb, cr, 0)), cache)
        out["oracle"] = {"quality": q2, "ssim": s2, "trace": tr2, "score_wall_s": round(time.perf_counter() - t0, 3)}
        out["same_quality"] = q == q2
        out["same_path"] = [m for m, _ in tr] == [m for m, _ in tr2]
        out["max_score_diff"] = max(abs(a[1] - b[1]) for a, b in zip(tr, tr2)) if out["same_path"] else None
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
