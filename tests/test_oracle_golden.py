"""CPU suite: the oracle against the frozen golden vectors, the independent NumPy restatement and the
reference's own (inequality) tests.  PARITY UNPINNED against the Go reference itself — no Go toolchain
exists here and the reference's tests hold no numeric vectors for this path."""
import hashlib
import math

import numpy as np
import pytest

from fennec_b200 import synth as S
from oracle import np_restatement as N
from tests import cases


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(cases.SCORE_CASES))
def test_oracle_scores_match_golden(name, golden, oracle):
    op, build = cases.SCORE_CASES[name]
    a, b = build()
    g = golden["scores"][name]
    assert [sha(a), sha(b)] == g["inputs_sha256"], "synthetic input generator drifted"
    got = getattr(oracle, op)(a, b)
    assert got == float.fromhex(g["hex"]), f"{name}: {got!r} vs golden {g['value']!r}"


@pytest.mark.parametrize("name", sorted(cases.PIXEL_CASES))
def test_oracle_pixels_match_golden(name, golden, golden_pixels, oracle):
    op, build, kw = cases.PIXEL_CASES[name]
    src = build()
    g = golden["pixels"][name]
    assert sha(src) == g["input_sha256"], "synthetic input generator drifted"
    out = getattr(oracle, op)(src, *kw.values())
    assert list(out.shape) == g["shape"]
    assert sha(out) == g["sha256"]
    if g["raw"]:
        assert np.array_equal(out, golden_pixels[name])


FAST_SCORE = ["ssim_r10_100", "ssim_small_4x4", "ssim_9x9_one_window", "ssim_ragged_131x77", "ssim_flat_250",
              "ssim_fast_small_300x200", "msssim_noise_100x60", "msssim_tiny_5x5", "msssim_noise_20x12"]


@pytest.mark.parametrize("name", FAST_SCORE)
def test_numpy_restatement_agrees_on_scores(name, golden):
    op, build = cases.SCORE_CASES[name]
    a, b = build()
    assert abs(getattr(N, op)(a, b, 8) - golden["scores"][name]["value"]) <= 1e-12


FAST_PIXEL = ["box_odd_ratio", "box_upsample", "blur_sigma2_noise", "blur_radius_gt_image", "blur3x3_noise",
              "sharpen_0p5_noise", "adaptive_0p3_stripes", "lanczos_down_4x", "lanczos_up", "lanczos_transparent"]


@pytest.mark.parametrize("name", FAST_PIXEL)
def test_numpy_restatement_agrees_on_pixels(name, golden):
    op, build, kw = cases.PIXEL_CASES[name]
    out = getattr(N, op)(build(), *kw.values())
    assert sha(out) == golden["pixels"][name]["sha256"]


def test_tables_match_golden(golden, oracle):
    k = oracle.gaussian_kernel(8, 1.5)
    assert sha(k) == golden["tables"]["ssim_kernel_8x8_sha256"]
    assert abs(k.sum() - 1.0) < 1e-15 and k[36] == golden["tables"]["ssim_kernel_center"]
    for key, (dst, src) in {"lanczos_7680_to_1920": (1920, 7680), "lanczos_400_to_100": (100, 400),
                            "lanczos_120_to_333": (333, 120)}.items():
        st, ix, wt = oracle.lanczos_weights(dst, src)
        assert sha(wt) == golden["tables"][key]["weights_sha256"]
        assert sha(ix.astype(np.int32)) == golden["tables"][key]["index_sha256"]


# ---- known answers and the reference's own inequality tests (fennec_test.go) on the oracle ----------

def test_black_white_is_analytic(oracle):
    # every window: mu = (0,255), sigma = 0  →  C1 / (255^2 + C1)
    v = oracle.ssim(S.make_solid_image(100, 100, (0, 0, 0, 255)), S.make_solid_image(100, 100, (255, 255, 255, 255)))
    assert abs(v - 6.5025 / (255.0 * 255.0 + 6.5025)) < 1e-15


def test_clampf_rounds_half_away_from_zero(oracle):
    for x, want in [(0.5, 1), (1.5, 2), (2.5, 3), (254.5, 255), (-0.5, 0), (0.49999999999999994, 0),
                    (-3.0, 0), (300.2, 255), (127.49999999, 127)]:
        assert oracle.clampf(x) == want


def test_reference_ssim_inequalities(oracle):  # fennec_test.go:82-129
    img = S.make_test_image(100, 100)
    assert oracle.ssim(img, img) >= 0.999
    assert oracle.ssim(S.make_solid_image(100, 100, (0, 0, 0, 255)), S.make_solid_image(100, 100, (255,) * 4)) <= 0.1
    assert 0.85 <= oracle.ssim(img, S.minus_red(img, 10)) <= 0.999
    big = S.make_test_image(500, 500)
    assert oracle.ssim_fast(big, big) >= 0.999
    small = S.make_test_image(4, 4)
    assert oracle.ssim(small, small) >= 0.999


def test_reference_msssim_inequalities(oracle):  # fennec_test.go:131-163
    img = S.make_test_image(128, 128)
    assert oracle.msssim(img, img) >= 0.99
    assert oracle.msssim(S.make_solid_image(128, 128, (0, 0, 0, 255)), S.make_solid_image(128, 128, (255,) * 4)) <= 0.1
    assert 0.7 <= oracle.msssim(img, S.minus_red(img, 5)) < 1.0


def test_reference_resize_and_effects_behaviour(oracle):  # fennec_test.go:510-560, 612-736, 1101-1115
    img = S.make_test_image(200, 100)
    assert oracle.lanczos_resize(img, 100, 50).shape == (50, 100, 4)
    assert oracle.lanczos_resize(img, 0, 50).shape == (0, 0, 4)
    rt = oracle.lanczos_resize(oracle.lanczos_resize(img, 100, 50), 200, 100)
    assert oracle.ssim(img, rt) >= 0.5
    assert oracle.smart_resize_dims(200, 100, 100, 100) == (False, 100, 50)
    assert oracle.smart_resize_dims(200, 100, 400, 400)[0] is True
    assert oracle.sharpen(img, 0.0) is img and oracle.adaptive_sharpen(img, 0.0) is img
    tiny = S.make_test_image(2, 2)
    assert oracle.sharpen(tiny, 0.5) is tiny and oracle.adaptive_sharpen(tiny, 0.5) is tiny
    assert oracle.gaussian_blur(img, 0.0) is img and oracle.gaussian_blur(img, -1.0) is img
    st = S.make_striped_image(100, 100, 10)
    assert np.any(oracle.sharpen(st, 0.8) != st)
    lin = S.make_test_image(100, 100)
    assert np.array_equal(oracle.sharpen(lin, 0.5), lin)  # a linear ramp is a fixed point of the 3x3 blur
    assert np.any(oracle.adaptive_sharpen(st, 0.5) != st)
    bl = oracle.gaussian_blur(img, 2.0)
    assert bl.shape == img.shape and oracle.ssim(img, bl) >= 0.3
    assert oracle.ssim(st, oracle.gaussian_blur(st, 20.0)) <= 0.999
    assert oracle.box_downsample(img, 50, 25).shape == (25, 50, 4)
    assert oracle.box_downsample(img, 0, 10).shape == (0, 0, 4)


def test_alpha_semantics(oracle):  # blur/sharpen pass alpha through, box averages it
    img = S.noise_image(64, 48, 5, alpha="random")
    for out in (oracle.gaussian_blur(img, 1.5), oracle.sharpen(img, 0.4), oracle.adaptive_sharpen(img, 0.4),
                oracle.blur3x3(img)):
        assert np.array_equal(out[..., 3], img[..., 3])
    assert not np.array_equal(oracle.box_downsample(img, 32, 24)[..., 3], img[::2, ::2, 3])
    assert np.array_equal(oracle.sharpen(img, 0.4)[0], img[0])  # Sharpen leaves border pixels unchanged
