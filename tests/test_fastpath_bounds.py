"""Host-side proof obligations of the FP32 fast paths (no GPU): a NumPy float32 emulation of the exact operation
sequence of AdaptiveSharpen's fast path (csrc/effects.cu, fx_tile_kernel MODE 2) must never disagree with the
oracle on a pixel it does not flag as ambiguous, and must flag only a small fraction."""
import numpy as np
import pytest

from fennec_b200 import synth as S

f32 = np.float32


def emulate_adaptive_fast(src: np.ndarray, strength: float):
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
    edge = np.minimum(np.sqrt(g2) * f32(2.5e-6), f32(1.0)).astype(f32)
    la = (f32(amount) * edge).astype(f32)
    out = np.zeros((h - 2, w - 2, 3), np.uint8)
    amb = np.zeros((h - 2, w - 2), bool)
    for ch in range(3):
        orig = P[1:h - 1, 1:w - 1, ch]
        diff = orig - bl[..., ch]
        t = (la * diff.astype(f32)).astype(f32)
        v = (orig.astype(f32) + t).astype(f32)
        lim = f32(0.5) - (np.abs(t).astype(np.float64) * np.float64(f32(5.3e-7)) + np.float64(f32(8e-5))).astype(f32)
        v = np.minimum(np.maximum(v, f32(-1)), f32(256))
        rounded = np.rint(v)
        amb |= np.abs(v - rounded) >= lim
        out[..., ch] = np.clip(rounded, 0, 255).astype(np.uint8)
    return out, amb


@pytest.mark.parametrize("strength", [0.5, 0.3, 0.77, 1.0])
@pytest.mark.parametrize("kind", ["noise", "photo", "stripes"])
def test_adaptive_fast_path_bound_is_sound(kind, strength, oracle):
    img = {"noise": lambda: S.noise_image(320, 240, 3, alpha="random"), "photo": lambda: S.gradient_noise_image(400, 300, 5),
           "stripes": lambda: S.make_striped_image(320, 200, 7)}[kind]()
    want = oracle.adaptive_sharpen(img, strength)[1:-1, 1:-1, :3]
    got, amb = emulate_adaptive_fast(img, strength)
    wrong = (got != want).any(-1)
    assert not (wrong & ~amb).any(), "a pixel outside the error bound disagrees with the reference arithmetic"
    assert amb.mean() < 0.01
