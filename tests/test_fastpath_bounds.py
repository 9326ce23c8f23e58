"""Host-side proof obligations of the FP32 fast paths (no GPU): a NumPy float32 emulation of the exact operation
sequence of AdaptiveSharpen's fast path (csrc/effects.cu, fx_tile_kernel MODE 2) must never disagree with the
oracle on a pixel it does not flag as ambiguous, and must flag only a small fraction."""
import numpy as np
import pytest

from fennec_b200 import synth as S

f32 = np.float32


def emulate_adaptive_fast(src: np.ndarray, strength: float, rsqrt_rel_err: float = 0.0):
    """csrc/effects.cu adaptive_tile_kernel, operation by operation in float32.  The device's rsqrt (MUFU, <= 2 ulp)
    cannot be reproduced bit for bit on the host, so |grad| is computed with a correctly rounded sqrt and then
    perturbed by `rsqrt_rel_err` (the tests sweep 0 and +-2.4e-7 = 2 ulp) — the bound must hold for all of them."""
    s = min(strength, 1.0)
    amount = 1.0 + 2.0 * s
    h, w, _ = src.shape
    P = src[..., :3].astype(np.int64)
    L = 299 * P[..., 0] + 587 * P[..., 1] + 114 * P[..., 2]                # exact integer lumas x1000
    k = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]])
    bl = np.zeros((h - 2, w - 2, 3), np.int64)
    for dy in range(3):
        for dx in range(3):
            bl += k[dy, dx] * P[dy:dy + h - 2, dx:dx + w - 2]
    bl = (bl + 8) >> 4                                                       # effects.go:124-136, exact
    sh = lambda a, dy, dx: a[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx]           # noqa: E731
    GX = -sh(L, -1, -1) + sh(L, -1, 1) - 2 * sh(L, 0, -1) + 2 * sh(L, 0, 1) - sh(L, 1, -1) + sh(L, 1, 1)
    GY = -sh(L, -1, -1) - 2 * sh(L, -1, 0) - sh(L, -1, 1) + sh(L, 1, -1) + 2 * sh(L, 1, 0) + sh(L, 1, 1)
    gxf, gyf = GX.astype(f32), GY.astype(f32)
    g2 = (gxf.astype(np.float64) * gxf.astype(np.float64) + (gyf * gyf).astype(np.float64)).astype(f32)   # fmaf
    g2 = np.maximum(g2, f32(1e-30))
    mag = (np.sqrt(g2.astype(np.float64)) * (1.0 + rsqrt_rel_err)).astype(f32)   # g2 * rsqrt(g2)
    edge = np.minimum((mag * f32(2.5e-6)).astype(f32), f32(1.0))
    la = (f32(amount) * edge).astype(f32)
    lim = f32(0.5) - (la.astype(np.float64) * np.float64(f32(255.0 * 7.5e-7)) + np.float64(f32(8e-5))).astype(f32)
    out = np.zeros((h - 2, w - 2, 3), np.uint8)
    worst = np.zeros((h - 2, w - 2), f32)
    for ch in range(3):
        orig = P[1:h - 1, 1:w - 1, ch]
        diff = (orig - bl[..., ch]).astype(f32)
        v = (la.astype(np.float64) * diff.astype(np.float64) + orig.astype(np.float64)).astype(f32)   # one rounding
        v = np.minimum(np.maximum(v, f32(0)), f32(255))
        rounded = np.rint(v)
        worst = np.maximum(worst, np.abs(v - rounded).astype(f32))
        out[..., ch] = rounded.astype(np.uint8)
    return out, worst >= lim


@pytest.mark.parametrize("strength", [0.5, 0.3, 0.77, 1.0])
@pytest.mark.parametrize("kind", ["noise", "photo", "stripes"])
def test_adaptive_fast_path_bound_is_sound(kind, strength, oracle):
    img = {"noise": lambda: S.noise_image(320, 240, 3, alpha="random"), "photo": lambda: S.gradient_noise_image(400, 300, 5),
           "stripes": lambda: S.make_striped_image(320, 200, 7)}[kind]()
    want = oracle.adaptive_sharpen(img, strength)[1:-1, 1:-1, :3]
    for err in (0.0, 2.4e-7, -2.4e-7):
        got, amb = emulate_adaptive_fast(img, strength, err)
        wrong = (got != want).any(-1)
        assert not (wrong & ~amb).any(), "a pixel outside the error bound disagrees with the reference arithmetic"
        assert amb.mean() < 0.01


# ---- round 2: the constant bounds of the Lanczos opaque-window shortcut and of the blur fast path ----------------------

def _fma32_chain(x: np.ndarray, w32: np.ndarray) -> np.ndarray:
    """acc = fmaf(x_k, w_k, acc) over the last axis in float32 (products of a byte and a float32 are exact in float64;
    the float64 sum is rounded once to float32: the double rounding is far below the margins tested here)."""
    acc = np.zeros(x.shape[:-1], f32)
    for k in range(x.shape[-1]):
        acc = (x[..., k].astype(np.float64) * np.float64(w32[k]) + acc.astype(np.float64)).astype(f32)
    return acc


def _windows(taps: int, w: np.ndarray, n: int, seed: int) -> np.ndarray:
    """Random byte windows plus the adversaries: extreme overshoot both ways, all-255, alternating, single spikes."""
    rng = np.random.default_rng(seed)
    adv = [np.where(w > 0, 255, 0), np.where(w > 0, 0, 255), np.full(taps, 255), np.full(taps, 1), np.arange(taps) % 2 * 255]
    adv += [np.eye(taps, dtype=np.int64)[k] * 255 for k in range(taps)]
    return np.concatenate([rng.integers(0, 256, (n, taps)), np.stack(adv)]).astype(np.int64)


def test_lanczos_opaque_shortcut_bound_is_sound(oracle):
    """csrc/api.cu detect_int_ratio: E = 255 * 2^-24 * (sum|wn| + sum_k P_k) * 1.05 + 1e-6 must bound the distance between the
    FP32 FMA chain over the normalised weights (csrc/resize.cu int_ratio_window, opaque branch) and the reference's
    binary64 sequence r += R * (255 w); a += 255 w; v = r * (1 / a)  (resize.go:99-110) — for the ratio-4 Lanczos-3 row."""
    start, index, weight = oracle.lanczos_weights(1920, 7680)
    mid = 960
    w = np.asarray(weight[start[mid]:start[mid + 1]], np.float64)
    assert len(w) == 24
    W = w.sum()
    wn = (w / W).astype(f32)
    P = np.cumsum(np.abs(wn.astype(np.float64)))
    Eo = 255.0 * 2.0 ** -24 * (P[-1] + P.sum()) * 1.05 + 1e-6
    assert 2.5e-4 < Eo < 3.5e-4                                   # what the kernel comment quotes (3.0e-4)
    x = _windows(24, w, 200_000, 7)
    r = np.zeros(len(x)); a = np.zeros(len(x))
    for k in range(24):                                           # the reference's order and operations, alpha = 255
        aw = 255.0 * w[k]
        r = r + x[:, k].astype(np.float64) * aw
        a = a + aw
    ref = r * (1.0 / a)
    got = _fma32_chain(x, wn).astype(np.float64)
    assert np.abs(got - ref).max() <= Eo
    assert np.abs(got - ref).max() > Eo / 50                      # ... and the bound is not vacuous
    # outputs the shortcut does NOT flag round like the reference
    amb = np.abs(got - np.rint(got)) >= 0.5 - Eo
    clamp = lambda v: np.clip(np.floor(v + 0.5), 0, 255)         # noqa: E731  (clampF: half away from zero, v > -0.5 here or clamped)
    assert (clamp(got)[~amb] == clamp(ref)[~amb]).all()
    assert amb.mean() < 0.003


@pytest.mark.parametrize("sigma", [0.8, 2.0, 2.66])
def test_blur_fast_path_bound_is_sound(sigma, oracle):
    """csrc/effects.cu launch_gaussian_blur: eps = 255 * 2^-24 * (sum|w32| + sum_k P_k) * 1.10 bounds the FP32 FMA chain against
    the reference's binary64 multiply-then-add sequence (effects.go:172-188)."""
    k64, radius = oracle.blur_kernel(sigma)
    k64 = np.asarray(k64, np.float64)
    w32 = k64.astype(f32)
    P = np.cumsum(np.abs(w32.astype(np.float64)))
    eps = 255.0 * 2.0 ** -24 * (P[-1] + P.sum()) * 1.10
    taps = 2 * radius + 1
    x = _windows(taps, k64 - k64.mean(), 200_000, int(sigma * 100))
    ref = np.zeros(len(x))
    for k in range(taps):
        ref = ref + x[:, k].astype(np.float64) * k64[k]
    got = _fma32_chain(x, w32).astype(np.float64)
    assert np.abs(got - ref).max() <= eps
    amb = np.abs(got - np.rint(got)) >= 0.5 - eps
    clamp = lambda v: np.clip(np.floor(v + 0.5), 0, 255)         # noqa: E731
    assert (clamp(got)[~amb] == clamp(ref)[~amb]).all()
    assert amb.mean() < 0.002
