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
