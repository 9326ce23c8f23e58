"""Analyze (analyze.go:26-176).

CPU part: the C oracle against the independent NumPy restatement and the reference's own TestAnalyze inequalities
(fennec_test.go:564-610).  GPU part: the CUDA scans through the C ABI — integer fields and EdgeDensity exact,
MeanBrightness / Contrast / Entropy within 1e-9 relative (the reference adds millions of doubles sequentially;
the device sums exact integers), recommendations identical.
"""
import numpy as np
import pytest

from fennec_b200 import synth as S

REL = 1e-9
FLOATS = ("entropy", "edge_density", "mean_brightness", "contrast", "estimated_compression")
INTS = ("width", "height", "has_alpha", "is_grayscale", "unique_colors", "recommended_format", "recommended_quality")

CASES = {
    "gradient_200": lambda: S.make_test_image(200, 200),
    "solid_gray_100": lambda: S.make_solid_image(100, 100, (128, 128, 128, 255)),
    "alpha_100": lambda: S.make_test_image_with_alpha(100, 100),
    "noise_640x480_alpha": lambda: S.noise_image(640, 480, 3, alpha="random"),
    "photo_1300x700": lambda: S.gradient_noise_image(1300, 700, 5),
    "stripes_333x217": lambda: S.make_striped_image(333, 217, 5),
    "tiny_2x2": lambda: S.noise_image(2, 2, 1),
    "thin_1x300": lambda: S.noise_image(1, 300, 2),
    "few_colours_500x400": lambda: (S.noise_image(500, 400, 4) & 0xC0) | 0x3F,       # 64 distinct colours, alpha 255
    "gray_noise_301x203": lambda: np.repeat(S.noise_image(301, 203, 6)[..., :1], 4, axis=2) | np.array([0, 0, 0, 255], np.uint8),
}


def _same(a, b, rel=REL):
    for k in INTS:
        assert a[k] == b[k], (k, a[k], b[k])
    for k in FLOATS:
        assert abs(a[k] - b[k]) <= rel * max(1.0, abs(b[k])), (k, a[k], b[k])


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_numpy_restatement(name, oracle):
    from oracle import np_restatement as N
    img = np.ascontiguousarray(CASES[name]())
    a, b = oracle.analyze(img), N.analyze(img)
    for k in ("has_alpha", "is_grayscale", "unique_colors"):
        assert a[k] == b[k], k
    for k in ("entropy", "edge_density", "mean_brightness", "contrast"):
        assert abs(a[k] - b[k]) <= 1e-12 * max(1.0, abs(a[k])), (k, a[k], b[k])
    assert np.array_equal(a["histogram"], b["histogram"])


def test_reference_analyze_tests_on_oracle(oracle):   # fennec_test.go:564-610
    st = oracle.analyze(S.make_test_image(200, 200))
    assert (st["width"], st["height"]) == (200, 200) and not st["has_alpha"] and st["entropy"] >= 1
    st = oracle.analyze(S.make_solid_image(100, 100, (128, 128, 128, 255)))
    assert st["is_grayscale"] and st["entropy"] <= 0.01
    st = oracle.analyze(S.make_test_image_with_alpha(100, 100))
    assert st["has_alpha"] and st["recommended_format"] == 2   # PNG
    st = oracle.analyze(np.zeros((0, 0, 4), np.uint8))
    assert (st["width"], st["height"]) == (0, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_analyze_matches_oracle(name, lib, oracle):
    from fennec_b200 import api
    img = np.ascontiguousarray(CASES[name]())
    _same(api.Analyze(img), oracle.analyze(img))


@pytest.mark.gpu
def test_gpu_reference_analyze_tests(lib):   # fennec_test.go:564-610 through the GPU path
    from fennec_b200 import api
    st = api.Analyze(S.make_test_image(200, 200))
    assert (st["width"], st["height"]) == (200, 200) and not st["has_alpha"] and st["entropy"] >= 1
    st = api.Analyze(S.make_solid_image(100, 100, (128, 128, 128, 255)))
    assert st["is_grayscale"] and st["entropy"] <= 0.01
    st = api.Analyze(S.make_test_image_with_alpha(100, 100))
    assert st["has_alpha"] and st["recommended_format"] == api.FORMAT_PNG
    st = api.Analyze(np.zeros((0, 0, 4), np.uint8))
    assert (st["width"], st["height"]) == (0, 0)


@pytest.mark.gpu
def test_gpu_analyze_batch_and_strided(lib, oracle):
    import torch
    from fennec_b200 import api, batch
    imgs = [S.gradient_noise_image(320, 240, 20 + i) for i in range(3)] + [S.noise_image(320, 240, 30, alpha="random")]
    got = batch.analyze_batch(torch.from_numpy(np.stack(imgs)).cuda())
    for g, img in zip(got, imgs):
        _same(g, oracle.analyze(img))
    wide = S.noise_image(400, 100, 40)
    view = wide[:, 7:306]                      # stride > 4*w, odd width → scalar tail + unaligned rows
    _same(api.Analyze(view), oracle.analyze(np.ascontiguousarray(view)))


@pytest.mark.gpu
def test_gpu_analyze_4k_vs_oracle(lib, oracle):
    from fennec_b200 import api
    img = S.gradient_noise_image(3840, 2160, 77)
    _same(api.Analyze(img), oracle.analyze(img))
