"""Device-resident timing of every op of the hot path at the BASELINE.json sizes (CUDA events, batch API).
Prints one JSON object per op: ms/step, work rate and achieved algorithmic GB/s vs the measured HBM peak."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fennec_b200 import batch
PEAK = 6533.8
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
only = sys.argv[1:] 

def noise(n, h, w, seed, alpha=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randint(0, 256, (n, h, w, 4), dtype=torch.uint8, device="cuda", generator=g)
    if not alpha: t[..., 3] = 255
    return t

def timeit(fn, iters):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def report(name, ms, n, mp_per_item, bytes_per_item, note=""):
    gbs = n * bytes_per_item / ms / 1e6
    print(json.dumps({"op": name, "n": n, "ms": round(ms, 4), "MP_per_s": round(n * mp_per_item / ms * 1e3, 1),
                      "items_per_s": round(n / ms * 1e3, 1), "algorithmic_GBps": round(gbs, 1),
                      "frac_of_measured_hbm": round(gbs / PEAK, 4), "note": note}), flush=True)

def want(k): return not only or k in only

if want("ssim"):
    a, b = noise(16, 2160, 3840, 1), noise(16, 2160, 3840, 2)
    report("SSIM 3840x2160 (metric)", timeit(lambda: batch.ssim_batch(a, b), 20), 16, 8.2944, 2 * 3840 * 2160 * 4)
    del a, b
if want("ssim_fast"):
    a, b = noise(16, 3024, 4032, 3), noise(16, 3024, 4032, 4)
    report("SSIMFast 4032x3024 (config 2)", timeit(lambda: batch.ssim_fast_batch(a, b), 20), 16, 12.192768, 2 * 4032 * 3024 * 4)
    del a, b
if want("blur"):
    x = noise(16, 2160, 3840, 5); y = torch.empty_like(x)
    report("GaussianBlur s=2 3840x2160", timeit(lambda: batch.gaussian_blur_batch(x, 2.0, out=y), 5), 16, 8.2944, 2 * 3840 * 2160 * 4)
    report("Sharpen 0.5 3840x2160", timeit(lambda: batch.sharpen_batch(x, 0.5, out=y), 10), 16, 8.2944, 2 * 3840 * 2160 * 4)
    report("AdaptiveSharpen 0.5 3840x2160", timeit(lambda: batch.adaptive_sharpen_batch(x, 0.5, out=y), 5), 16, 8.2944, 2 * 3840 * 2160 * 4)
    z = torch.empty_like(x)
    def both():
        batch.gaussian_blur_batch(x, 2.0, out=y); batch.sharpen_batch(y, 0.5, out=z)
    report("Blur s=2 + Sharpen 0.5 (config 3)", timeit(both, 5), 16, 8.2944, 4 * 3840 * 2160 * 4)
    del x, y, z
if want("lanczos"):
    x = noise(8, 4320, 7680, 6); y = torch.zeros((8, 1080, 1920, 4), dtype=torch.uint8, device="cuda")
    report("Lanczos3 7680x4320->1920x1080 (config 4)", timeit(lambda: batch.lanczos_resize_batch(x, 1920, 1080, out=y), 3), 8, 33.1776,
           7680 * 4320 * 4 + 1920 * 1080 * 4)
    x[..., 3] = 255
    report("Lanczos3 7680x4320->1920x1080 opaque (config 4)", timeit(lambda: batch.lanczos_resize_batch(x, 1920, 1080, out=y), 3), 8, 33.1776,
           7680 * 4320 * 4 + 1920 * 1080 * 4)
    del x, y
if want("ycbcr"):
    # convertToNRGBA of a decoded 4:2:0 JPEG, device-resident (1.5 B read + 4 B written per pixel)
    n, h, w = 16, 3024, 4032
    g = torch.Generator(device="cuda").manual_seed(11)
    ty = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device="cuda", generator=g)
    tcb = torch.randint(0, 256, (n, h // 2, w // 2), dtype=torch.uint8, device="cuda", generator=g)
    tcr = torch.randint(0, 256, (n, h // 2, w // 2), dtype=torch.uint8, device="cuda", generator=g)
    out = torch.empty((n, h, w, 4), dtype=torch.uint8, device="cuda")
    report("convertToNRGBA 4:2:0 4032x3024 (f1)", timeit(lambda: batch.ycbcr_to_nrgba_batch(ty, tcb, tcr, 2, out=out), 20), n, 12.192768,
           int(4032 * 3024 * 5.5))
    del ty, tcb, tcr, out
    # one iteration of the quality search (compress.go:45-74), host buffers, PCIe inside the timed region:
    # reference semantics (both NRGBA images uploaded, src downsampled again) vs the cached-source session + YCbCr upload
    import time
    import numpy as np
    from fennec_b200 import api, synth
    src = synth.gradient_noise_image(4032, 3024, 5)
    py, pcb, pcr = synth.ycbcr_planes_from_nrgba(src, 2, 1, 3)
    cand = api.ycbcr_to_nrgba(py, pcb, pcr, 2)
    def wall(fn, iters=10):
        fn(); t0 = time.perf_counter()
        for _ in range(iters): fn()
        return (time.perf_counter() - t0) / iters * 1e3
    ms_ref = wall(lambda: api.SSIMFast(src, cand))
    with api.SSIMReference(src) as ref:
        ms_ses = wall(lambda: ref.score_ycbcr(py, pcb, pcr, 2))
        ms_ses_n = wall(lambda: ref.score_nrgba(cand))
    print(json.dumps({"op": "search iteration e2e 4032x3024 (config 2), pageable host buffers", "ms_fb_ssim_fast_both_nrgba": round(ms_ref, 3),
                      "ms_session_ycbcr420": round(ms_ses, 3), "ms_session_nrgba": round(ms_ses_n, 3),
                      "h2d_bytes": {"fb_ssim_fast": 2 * 4032 * 3024 * 4, "session_ycbcr420": int(4032 * 3024 * 1.5), "session_nrgba": 4032 * 3024 * 4}}), flush=True)
if want("analyze"):
    # the scans behind Analyze, device-resident (4 B/px read once + sampled reads)
    x = noise(16, 2160, 3840, 12)
    from fennec_b200 import _lib
    raw = torch.empty(16 * int(_lib.load().fb_analyze_raw_bytes()), dtype=torch.uint8, device="cuda")
    report("Analyze scans 3840x2160 (f2)", timeit(lambda: batch.analyze_scan_batch(x, raw), 20), 16, 8.2944, 3840 * 2160 * 4)
    del x, raw
if want("orient"):
    x = noise(16, 2160, 3840, 13)
    y = torch.empty((16, 3840, 2160, 4), dtype=torch.uint8, device="cuda")
    z = torch.empty_like(x)
    report("ApplyOrientation Rotate90CW 3840x2160 (f4)", timeit(lambda: batch.apply_orientation_batch(x, 6, out=y), 20), 16, 8.2944, 2 * 3840 * 2160 * 4)
    report("ApplyOrientation FlipH 3840x2160 (f4)", timeit(lambda: batch.apply_orientation_batch(x, 2, out=z), 20), 16, 8.2944, 2 * 3840 * 2160 * 4)
    del x, y, z
if want("palette"):
    x = noise(4, 3024, 4032, 14)
    pal = torch.randint(0, 256, (4, 256, 4), dtype=torch.uint8, device="cuda")
    pal[..., 3] = 255
    report("applyPalette 256 colours 4032x3024, noise image + random palette (f3)", timeit(lambda: batch.apply_palette_batch(x, pal, 256), 5), 4, 12.192768,
           4032 * 3024 * 9)
    # photo-like: smooth gradients + noise sigma 6; palette = 256 colours sampled from the image (what median cut yields)
    yy = torch.arange(3024, device="cuda").view(1, -1, 1, 1) / 3024.0
    xx = torch.arange(4032, device="cuda").view(1, 1, -1, 1) / 4032.0
    k = torch.tensor([[1.0, 0.2], [0.3, 0.9], [0.6, 0.6], [0.0, 0.0]], device="cuda")
    base = 230.0 * (xx * k[:, 0].view(1, 1, 1, 4) + yy * k[:, 1].view(1, 1, 1, 4)) / (k.sum(1).view(1, 1, 1, 4) + 1e-6)
    g = torch.Generator(device="cuda").manual_seed(15)
    ph = (base + 6.0 * torch.randn((4, 3024, 4032, 4), device="cuda", generator=g)).clamp(0, 255).to(torch.uint8)
    ph[..., 3] = 255
    flat = ph.view(4, -1, 4)
    pick = torch.randint(0, flat.shape[1], (256,), device="cuda", generator=g)
    pal2 = flat[:, pick].contiguous()
    report("applyPalette 256 colours 4032x3024, photo-like image + palette sampled from it (f3)", timeit(lambda: batch.apply_palette_batch(ph, pal2, 256), 5), 4,
           12.192768, 4032 * 3024 * 9)
    del x, pal, ph, pal2, flat, base
if want("pixfmt"):
    x = noise(16, 3024, 4032, 21)
    y = torch.empty_like(x)
    x[..., 3] = 255
    report("convertToNRGBA *image.RGBA opaque 4032x3024", timeit(lambda: batch.convert_to_nrgba_batch(1, x, out=y), 10), 16, 12.192768, 4032 * 3024 * 8)
    x = noise(16, 3024, 4032, 22)
    x[..., :3] = (x[..., :3].to(torch.int32) * x[..., 3:4].to(torch.int32) // 255).to(torch.uint8)
    report("convertToNRGBA *image.RGBA translucent (3 divisions per pixel) 4032x3024", timeit(lambda: batch.convert_to_nrgba_batch(1, x, out=y), 10), 16, 12.192768, 4032 * 3024 * 8)
    x64 = torch.randint(0, 256, (8, 3024, 4032, 8), dtype=torch.uint8, device="cuda")
    y8 = torch.empty((8, 3024, 4032, 4), dtype=torch.uint8, device="cuda")
    report("convertToNRGBA *image.NRGBA64 random alpha 4032x3024", timeit(lambda: batch.convert_to_nrgba_batch(3, x64, out=y8), 10), 8, 12.192768, 4032 * 3024 * 12)
    del x, y, x64, y8
if want("msssim"):
    a, b = noise(16, 4320, 7680, 7), noise(16, 4320, 7680, 8)
    report("MSSSIM 7680x4320 (config 5)", timeit(lambda: batch.msssim_batch(a, b), 5), 16, 33.1776, 2 * 7680 * 4320 * 4)
    del a, b
    x = noise(8, 4320, 7680, 9)
    y = torch.zeros((8, 288, 512, 4), dtype=torch.uint8, device="cuda")
    report("boxDownsample 7680x4320->512x288", timeit(lambda: batch.box_downsample_batch(x, 512, 288, out=y), 10), 8, 33.1776, 7680 * 4320 * 4)
    y2 = torch.zeros((8, 2160, 3840, 4), dtype=torch.uint8, device="cuda")
    report("boxDownsample 7680x4320->3840x2160", timeit(lambda: batch.box_downsample_batch(x, 3840, 2160, out=y2), 10), 8, 33.1776, 7680 * 4320 * 4 * 1.25)
