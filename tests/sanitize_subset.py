"""A compact walk over every kernel family with awkward shapes (ragged widths, unaligned strides, tiny images),
meant to run under `compute-sanitizer --tool memcheck`. Also checks results against the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
from fennec_b200 import api, synth as S
from oracle import pyoracle as O

def padded(img, pad):
    h, w = img.shape[:2]
    buf = np.full((h, w + pad, 4), 0xCD, dtype=np.uint8); buf[:, :w] = img
    return buf[:, :w]

for (w, h) in [(9, 9), (13, 40), (131, 77), (248, 64), (250, 33), (512, 20), (641, 19)]:
    a = S.noise_image(w, h, w + h, alpha="random"); b = S.perturb(a, 7, 9)
    for aa, bb in ((a, b), (padded(a, 3), padded(b, 5))):
        assert abs(api.SSIM(aa, bb) - O.ssim(a, b)) <= 1e-5
        assert abs(api.SSIMFast(aa, bb) - O.ssim_fast(a, b)) <= 1e-5
        assert abs(api.MSSSIM(aa, bb) - O.msssim(a, b)) <= 1e-5
    for src in (a, padded(a, 1)):
        assert np.array_equal(api.GaussianBlur(src, 2.0), O.gaussian_blur(a, 2.0))
        assert np.array_equal(api.GaussianBlur(src, 0.7), O.gaussian_blur(a, 0.7))
        assert np.array_equal(api.GaussianBlur(src, 4.0), O.gaussian_blur(a, 4.0))
        assert np.array_equal(api.Sharpen(src, 0.5), O.sharpen(a, 0.5))
        assert np.array_equal(api.Sharpen(src, 0.3), O.sharpen(a, 0.3))
        assert np.array_equal(api.AdaptiveSharpen(src, 0.6), O.adaptive_sharpen(a, 0.6))
        assert np.array_equal(api.blur3x3(src), O.blur3x3(a))
        for (dw, dh) in ((max(1, w // 3), max(1, h // 2)), (w + 5, h + 3), (1, 1)):
            assert np.array_equal(api.lanczos_resize(src, dw, dh), O.lanczos_resize(a, dw, dh))
            assert np.array_equal(api.box_downsample(src, dw, dh), O.box_downsample(a, dw, dh))
big = S.noise_image(2051, 517, 3, alpha="random")
assert np.array_equal(api.box_downsample(big, 512, 129), O.box_downsample(big, 512, 129))
assert np.array_equal(api.lanczos_resize(big, 513, 130), O.lanczos_resize(big, 513, 130))
assert np.array_equal(api.GaussianBlur(big, 2.0), O.gaussian_blur(big, 2.0))
assert abs(api.SSIMFast(big, S.perturb(big, 1, 5)) - O.ssim_fast(big, S.perturb(big, 1, 5))) <= 1e-5
# round-1b kernels: fused MS-SSIM level step, integer-ratio Lanczos (pipelined H pass + exact queue), tall blur segments,
# YCbCr / Gray conversion, Analyze scans, orientation, palette
import torch
from fennec_b200 import batch
for (w, h) in [(1028, 770), (1300, 700), (2048, 1152)]:
    a = S.gradient_noise_image(w, h, w); b = S.perturb(a, 3, 8)
    assert abs(api.MSSSIM(a, b) - O.msssim(a, b)) <= 1e-5
    ok, tw, th = api.ssim_fast_dims(w, h)
    ta, tb, ha, hb = batch.msssim_level_batch(torch.from_numpy(a[None]).cuda(), torch.from_numpy(b[None]).cuda(), tw, th)
    assert np.array_equal(ta[0].cpu().numpy(), O.box_downsample(a, tw, th)) and np.array_equal(hb[0].cpu().numpy(), O.box_downsample(b, w // 2, h // 2))
for (w, h, al) in [(1024, 64, "opaque"), (2052, 40, "random"), (516, 36, "ramp")]:
    src = S.noise_image(w, h, w + 1, alpha=al)
    assert np.array_equal(api.lanczos_resize(src, w // 4, h // 4), O.lanczos_resize(src, w // 4, h // 4))
tall = S.noise_image(70, 500, 5, alpha="random")
assert np.array_equal(api.GaussianBlur(tall, 2.0), O.gaussian_blur(tall, 2.0))
for ratio in range(6):
    for (w, h) in [(67, 45), (1, 1), (130, 3)]:
        y, cb, cr = S.noise_planes(w, h, ratio, 10 + ratio)
        assert np.array_equal(api.ycbcr_to_nrgba(y, cb, cr, ratio), O.ycbcr_to_nrgba(y, cb, cr, ratio))
g = S.noise_image(37, 21, 5)[..., 0].copy()
assert np.array_equal(api.gray_to_nrgba(g), O.gray_to_nrgba(g))
src = S.gradient_noise_image(640, 480, 9)
with api.SSIMReference(src) as ref:
    y, cb, cr = S.ycbcr_planes_from_nrgba(src, 2, 1, 3)
    assert abs(ref.score_ycbcr(y, cb, cr, 2) - O.ssim_fast(src, O.ycbcr_to_nrgba(y, cb, cr, 2))) <= 1e-5
for img in (S.noise_image(301, 203, 6, alpha="random"), S.gradient_noise_image(640, 480, 7), S.noise_image(3, 3, 1)):
    st, want = api.Analyze(img), O.analyze(img)
    assert st["unique_colors"] == want["unique_colors"] and st["has_alpha"] == want["has_alpha"]
    assert abs(st["edge_density"] - want["edge_density"]) <= 1e-12 and abs(st["entropy"] - want["entropy"]) <= 1e-9
for o in range(2, 9):
    for (w, h) in [(33, 65), (1, 7), (128, 64), (257, 31)]:
        img = S.noise_image(w, h, w + o, alpha="random")
        assert np.array_equal(api.ApplyOrientation(img, o), O.apply_orientation(img, o))
rng = np.random.Generator(np.random.PCG64(3))
pal = rng.integers(0, 256, (200, 4), dtype=np.uint8); pal[:, 3] = 255
for (w, h) in [(97, 61), (4, 1), (130, 17)]:
    img = S.noise_image(w, h, w, alpha="random")
    gi, go = api.apply_palette(img, pal); oi, oo = O.apply_palette(img, pal)
    assert np.array_equal(gi, oi) and np.array_equal(go, oo)
# round-2 kernels: lean interior tiles of Sharpen / AdaptiveSharpen (need whole warps inside the image), warp-autonomous
# Lanczos H / V passes with the opaque shortcut and ragged widths, cp.async-staged blur V pass over several segments
wide = S.noise_image(700, 300, 11, alpha="random")
assert np.array_equal(api.Sharpen(wide, 0.5), O.sharpen(wide, 0.5))
assert np.array_equal(api.AdaptiveSharpen(wide, 0.6), O.adaptive_sharpen(wide, 0.6))
assert np.array_equal(api.GaussianBlur(wide, 2.0), O.gaussian_blur(wide, 2.0))
for (w, h, al) in [(1000, 64, "opaque"), (4 * 131, 40, "opaque"), (1536, 72, "random")]:
    src = S.noise_image(w, h, w + 2, alpha=al)
    assert np.array_equal(api.lanczos_resize(src, w // 4, h // 4), O.lanczos_resize(src, w // 4, h // 4))
a8 = S.gradient_noise_image(1920, 1080, 13); b8 = S.perturb(a8, 4, 7)
assert abs(api.MSSSIM(a8, b8) - O.msssim(a8, b8)) <= 1e-5
print("sanitize subset ok")
