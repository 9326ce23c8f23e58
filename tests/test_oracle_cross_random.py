"""The reference has no golden vectors and no Go toolchain exists here, so the oracle is pinned by AGREEMENT between
two restatements written independently from the Go source — C (oracle/fennec_oracle.c, the checker) and vectorised
NumPy (oracle/np_restatement.py).  tests/test_oracle_golden.py compares them on the frozen cases; this file sweeps
seeded random shapes, ratios and parameters so that an index or rounding slip in either shows up off the golden set:
ragged dims, up- and down-scaling at non-integer ratios, sigma from sub-pixel to wider than the image, translucent
alpha, images below the 8-pixel SSIM window."""
import numpy as np
import pytest

from fennec_b200 import synth as S
from oracle import np_restatement as N

RNG = np.random.Generator(np.random.PCG64(20251017))
DIMS = [(int(w), int(h)) for w, h in RNG.integers(1, 70, (10, 2))] + [(1, 1), (8, 8), (9, 64), (64, 9), (7, 40)]


def _img(w, h, seed, alpha="random"):
    return np.ascontiguousarray(S.noise_image(w, h, seed, alpha=alpha))


@pytest.mark.parametrize("w,h", DIMS)
def test_box_downsample_random_targets(w, h, oracle):
    img = _img(w, h, w * 131 + h)
    r = np.random.Generator(np.random.PCG64(w * 7 + h))
    for _ in range(4):
        dw, dh = int(r.integers(1, 2 * w + 2)), int(r.integers(1, 2 * h + 2))
        assert np.array_equal(oracle.box_downsample(img, dw, dh), N.box_downsample(img, dw, dh)), (w, h, dw, dh)


@pytest.mark.parametrize("w,h", DIMS)
def test_lanczos_resize_random_targets(w, h, oracle):
    img = _img(w, h, w * 17 + h * 3)
    r = np.random.Generator(np.random.PCG64(w * 11 + h))
    for _ in range(3):
        dw, dh = int(r.integers(1, 2 * w + 3)), int(r.integers(1, 2 * h + 3))
        assert np.array_equal(oracle.lanczos_resize(img, dw, dh), N.lanczos_resize(img, dw, dh)), (w, h, dw, dh)


def test_lanczos_weight_tables_random_sizes(oracle):
    r = np.random.Generator(np.random.PCG64(5))
    for _ in range(60):
        src, dst = int(r.integers(1, 400)), int(r.integers(1, 400))
        start, index, weight = oracle.lanczos_weights(dst, src)          # CSR: taps of d at [start[d], start[d+1])
        table = N.lanczos_weights(dst, src)                              # list of (indices, weights) per d
        assert len(table) == dst
        for d, (idx, wts) in enumerate(table):
            lo, hi = int(start[d]), int(start[d + 1])
            assert index[lo:hi].tolist() == idx, (src, dst, d)
            assert np.max(np.abs(weight[lo:hi] - np.asarray(wts)), initial=0.0) <= 1e-15, (src, dst, d)


@pytest.mark.parametrize("w,h", DIMS)
def test_effects_random_parameters(w, h, oracle):
    img = _img(w, h, w * 5 + h * 9)
    r = np.random.Generator(np.random.PCG64(w + 100 * h))
    for sigma in (float(r.uniform(0.05, 1.0)), float(r.uniform(1.0, 4.0)), float(r.uniform(4.0, 12.0))):
        assert np.array_equal(oracle.gaussian_blur(img, sigma), N.gaussian_blur(img, sigma)), (w, h, sigma)
    for strength in (float(r.uniform(0.01, 1.0)), float(r.uniform(1.0, 3.0))):
        a, b = oracle.sharpen(img, strength), N.sharpen(img, strength)
        assert np.array_equal(a, b), (w, h, strength)
        a, b = oracle.adaptive_sharpen(img, strength), N.adaptive_sharpen(img, strength)
        assert np.array_equal(a, b), (w, h, strength)
    assert np.array_equal(oracle.blur3x3(img), N.blur3x3(img))


@pytest.mark.parametrize("w,h", DIMS)
def test_ssim_family_random_shapes(w, h, oracle):
    a = _img(w, h, w * 3 + h * 29)
    b = S.perturb(a, w + h, 9)
    assert abs(oracle.ssim(a, b) - N.ssim(a, b, 8)) <= 1e-12
    assert abs(oracle.pixel_ssim(a, b) - N.pixel_ssim(a, b)) <= 1e-12
    assert abs(oracle.msssim(a, b) - N.msssim(a, b, 8)) <= 1e-12


def test_ssim_fast_above_512_random(oracle):
    r = np.random.Generator(np.random.PCG64(9))
    for _ in range(3):
        w, h = int(r.integers(513, 900)), int(r.integers(20, 700))
        a = _img(w, h, w + h)
        b = S.perturb(a, w, 7)
        assert oracle.ssim_fast_dims(w, h) == N.ssim_fast_dims(w, h)
        assert abs(oracle.ssim_fast(a, b) - N.ssim_fast(a, b, 8)) <= 1e-12


def test_smart_resize_dims_random(oracle):
    r = np.random.Generator(np.random.PCG64(13))
    for _ in range(300):
        sw, sh, mw, mh = (int(v) for v in r.integers(1, 9000, 4))
        assert oracle.smart_resize_dims(sw, sh, mw, mh) == N.smart_resize_dims(sw, sh, mw, mh), (sw, sh, mw, mh)
