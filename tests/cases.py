"""Parity case registry shared by oracle/gen_golden.py and the tests.

Each case names an op of the hot path, the synthetic inputs (fennec_b200.synth — the reference's
own test generators plus seeded distributions) and the op's parameters.  Sizes are small enough
for the CPU oracle to finish in seconds.  Expected values live in tests/golden/golden.json.
"""
from __future__ import annotations

import numpy as np

from fennec_b200 import synth as S


def _pair_r10(w, h, delta=10):
    a = S.make_test_image(w, h)
    return a, S.minus_red(a, delta)


def _pair_noise(w, h, seed, amp=6):
    a = S.noise_image(w, h, seed)
    return a, S.perturb(a, seed + 1, amp)


def _pair_flat(w, h, base, seed):
    return S.flat_pm1_image(w, h, base, seed), S.flat_pm1_image(w, h, base, seed + 100)


def _pair_grad(w, h, seed):
    a = S.gradient_noise_image(w, h, seed)
    return a, S.perturb(a, seed + 1, 6)


def _padded(img: np.ndarray, pad_px: int) -> np.ndarray:
    """A view whose row stride is wider than 4*w (exercises Stride handling, ssim.go:212-216)."""
    h, w = img.shape[:2]
    buf = np.full((h, w + pad_px, 4), 0xAB, dtype=np.uint8)
    buf[:, :w] = img
    return buf[:, :w]


# name -> (op, builder returning the positional inputs, kwargs)
SCORE_CASES = {
    # the reference's own test inputs (fennec_test.go:82-163)
    "ssim_identical_100": ("ssim", lambda: (S.make_test_image(100, 100),) * 2),
    "ssim_black_white_100": ("ssim", lambda: (S.make_solid_image(100, 100, (0, 0, 0, 255)),
                                              S.make_solid_image(100, 100, (255, 255, 255, 255)))),
    "ssim_r10_100": ("ssim", lambda: _pair_r10(100, 100)),
    "ssim_small_4x4": ("ssim", lambda: _pair_r10(4, 4)),
    "ssim_7x20_pixel_path": ("ssim", lambda: _pair_noise(7, 20, 5)),
    "ssim_8x8_no_windows": ("ssim", lambda: _pair_noise(8, 8, 6)),
    "ssim_8x40_no_windows": ("ssim", lambda: _pair_noise(8, 40, 7)),
    "ssim_9x9_one_window": ("ssim", lambda: _pair_noise(9, 9, 8)),
    # BASELINE.json configs[0]
    "ssim_r10_640x480": ("ssim", lambda: _pair_r10(640, 480)),
    "ssim_noise_640x480": ("ssim", lambda: _pair_noise(640, 480, 1234)),
    "ssim_grad_640x480": ("ssim", lambda: _pair_grad(640, 480, 21)),
    "ssim_ragged_131x77": ("ssim", lambda: _pair_noise(131, 77, 9)),
    "ssim_ragged_257x35": ("ssim", lambda: _pair_grad(257, 35, 10)),
    "ssim_padded_stride": ("ssim", lambda: tuple(_padded(x, 5) for x in _pair_noise(150, 90, 11))),
    "ssim_stripes": ("ssim", lambda: (S.make_striped_image(200, 120, 10), S.make_striped_image(200, 120, 12))),
    # FP32 cancellation adversaries
    **{f"ssim_flat_{b}": ("ssim", (lambda b=b: _pair_flat(320, 200, b, 40 + b))) for b in (0, 3, 64, 128, 250, 254)},
    "ssim_checker_flat": ("ssim", lambda: (S.checker_flat_image(320, 200, 16, 2, 250, 50),
                                           S.checker_flat_image(320, 200, 16, 2, 250, 51))),
    "ssim_fast_500_identical": ("ssim_fast", lambda: (S.make_test_image(500, 500),) * 2),
    "ssim_fast_small_300x200": ("ssim_fast", lambda: _pair_noise(300, 200, 12)),
    "ssim_fast_1300x700": ("ssim_fast", lambda: _pair_grad(1300, 700, 5)),
    "ssim_fast_2016x1512": ("ssim_fast", lambda: _pair_grad(2016, 1512, 13)),
    "ssim_fast_600x9": ("ssim_fast", lambda: _pair_noise(600, 9, 14)),
    "ssim_fast_5000x40": ("ssim_fast", lambda: _pair_noise(5000, 40, 15)),
    "msssim_identical_128": ("msssim", lambda: (S.make_test_image(128, 128),) * 2),
    "msssim_black_white_128": ("msssim", lambda: (S.make_solid_image(128, 128, (0, 0, 0, 255)),
                                                  S.make_solid_image(128, 128, (255, 255, 255, 255)))),
    "msssim_r5_128": ("msssim", lambda: _pair_r10(128, 128, 5)),
    "msssim_noise_20x12": ("msssim", lambda: _pair_noise(20, 12, 3)),
    "msssim_noise_100x60": ("msssim", lambda: _pair_noise(100, 60, 16)),
    "msssim_tiny_5x5": ("msssim", lambda: _pair_noise(5, 5, 17)),
    "msssim_grad_1300x700": ("msssim", lambda: _pair_grad(1300, 700, 5)),
    "msssim_grad_1920x1080": ("msssim", lambda: _pair_grad(1920, 1080, 18)),
}

_ALPHA = lambda w, h, seed: S.noise_image(w, h, seed, alpha="random")  # noqa: E731

# name -> (op, builder returning the source image, kwargs)
PIXEL_CASES = {
    "box_1300x700_to_512x276": ("box_downsample", lambda: S.gradient_noise_image(1300, 700, 5), dict(dw=512, dh=276)),
    "box_odd_ratio": ("box_downsample", lambda: _ALPHA(1000, 333, 19), dict(dw=333, dh=111)),
    "box_half": ("box_downsample", lambda: _ALPHA(642, 481, 20), dict(dw=321, dh=240)),
    "box_upsample": ("box_downsample", lambda: _ALPHA(60, 50, 21), dict(dw=100, dh=120)),
    "box_to_1x1": ("box_downsample", lambda: _ALPHA(37, 23, 22), dict(dw=1, dh=1)),
    "box_ratio_7p875": ("box_downsample", lambda: _ALPHA(2016, 378, 23), dict(dw=256, dh=48)),
    "blur_sigma2_noise": ("gaussian_blur", lambda: _ALPHA(301, 203, 8), dict(sigma=2.0)),
    "blur_sigma0p5": ("gaussian_blur", lambda: _ALPHA(301, 203, 8), dict(sigma=0.5)),
    "blur_sigma3p3_grad": ("gaussian_blur", lambda: S.make_test_image(257, 131), dict(sigma=3.3)),
    "blur_sigma20_small": ("gaussian_blur", lambda: _ALPHA(90, 70, 24), dict(sigma=20.0)),
    "blur_radius_gt_image": ("gaussian_blur", lambda: _ALPHA(5, 3, 25), dict(sigma=2.0)),
    "blur_1x1": ("gaussian_blur", lambda: _ALPHA(1, 1, 26), dict(sigma=1.0)),
    "blur3x3_noise": ("blur3x3", lambda: _ALPHA(130, 67, 27), dict()),
    "sharpen_0p5_noise": ("sharpen", lambda: _ALPHA(301, 203, 8), dict(strength=0.5)),
    "sharpen_0p3_grad": ("sharpen", lambda: S.make_test_image(200, 200), dict(strength=0.3)),
    "sharpen_clamped_2p0": ("sharpen", lambda: _ALPHA(64, 48, 28), dict(strength=2.0)),
    "sharpen_3x3": ("sharpen", lambda: _ALPHA(3, 3, 29), dict(strength=0.7)),
    "adaptive_0p5_noise": ("adaptive_sharpen", lambda: _ALPHA(301, 203, 8), dict(strength=0.5)),
    "adaptive_0p3_stripes": ("adaptive_sharpen", lambda: S.make_striped_image(200, 120, 10), dict(strength=0.3)),
    "adaptive_1p0_grad": ("adaptive_sharpen", lambda: S.gradient_noise_image(160, 90, 30), dict(strength=1.0)),
    "lanczos_down_4x": ("lanczos_resize", lambda: S.noise_image(400, 300, 11, alpha="ramp"), dict(dw=100, dh=75)),
    "lanczos_down_odd": ("lanczos_resize", lambda: S.noise_image(400, 300, 11, alpha="ramp"), dict(dw=37, dh=299)),
    "lanczos_up": ("lanczos_resize", lambda: S.noise_image(120, 90, 31, alpha="random"), dict(dw=333, dh=200)),
    "lanczos_mixed": ("lanczos_resize", lambda: S.gradient_noise_image(400, 300, 32), dict(dw=555, dh=100)),
    "lanczos_alpha_ramp_200_to_50": ("lanczos_resize", lambda: S.make_test_image_with_alpha(200, 200), dict(dw=50, dh=50)),
    "lanczos_same_size_copy": ("lanczos_resize", lambda: _ALPHA(33, 21, 33), dict(dw=33, dh=21)),
    "lanczos_transparent": ("lanczos_resize", lambda: _transparent(64, 64, 34), dict(dw=20, dh=24)),
    "lanczos_h_only": ("lanczos_resize", lambda: _ALPHA(200, 50, 35), dict(dw=80, dh=50)),
    "lanczos_to_1x1": ("lanczos_resize", lambda: _ALPHA(40, 30, 36), dict(dw=1, dh=1)),
}


def _transparent(w, h, seed):
    """Mostly alpha==0 with a few opaque pixels: exercises the a<=0.5 'leave zero' branch."""
    img = S.noise_image(w, h, seed, alpha="random")
    mask = (np.arange(w)[None, :] + np.arange(h)[:, None]) % 5 != 0
    img[mask, 3] = 0
    return img
