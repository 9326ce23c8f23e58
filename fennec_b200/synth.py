"""Deterministic synthetic NRGBA inputs (numpy, host side).

The first four generators restate the reference's own test-image helpers
(fennec_test.go:20-76) so parity tests can use the inputs the reference's tests use; the rest
are the seeded distributions SURVEY.md §8d asks for (uniform noise, gradient + noise, flat±1
adversaries for FP32 cancellation, translucent ramps).  All return uint8 arrays (h, w, 4).
"""
from __future__ import annotations

import numpy as np


def make_test_image(w: int, h: int) -> np.ndarray:
    """fennec_test.go:20-32 — R=x*255/w, G=y*255/h, B=(x+y)%256, A=255 (integer division)."""
    x = np.arange(w, dtype=np.int64)[None, :]
    y = np.arange(h, dtype=np.int64)[:, None]
    img = np.empty((h, w, 4), dtype=np.uint8)
    img[..., 0] = np.broadcast_to((x * 255 // max(w, 1)) & 0xFF, (h, w))
    img[..., 1] = np.broadcast_to((y * 255 // max(h, 1)) & 0xFF, (h, w))
    img[..., 2] = (x + y) % 256
    img[..., 3] = 255
    return img


def make_test_image_with_alpha(w: int, h: int) -> np.ndarray:
    """fennec_test.go:34-43 — alpha ramps with x."""
    img = make_test_image(w, h)
    x = np.arange(w, dtype=np.int64)[None, :]
    img[..., 3] = np.broadcast_to((x * 255 // max(w, 1)) & 0xFF, (h, w))
    return img


def make_solid_image(w: int, h: int, rgba) -> np.ndarray:
    """fennec_test.go:45-54"""
    img = np.empty((h, w, 4), dtype=np.uint8)
    img[...] = np.asarray(rgba, dtype=np.uint8)
    return img


def make_striped_image(w: int, h: int, stripe_width: int) -> np.ndarray:
    """fennec_test.go:58-76 — alternating vertical stripes (200,50,100)/(50,200,100)."""
    x = np.arange(w)[None, :]
    even = ((x // stripe_width) % 2 == 0)
    img = np.empty((h, w, 4), dtype=np.uint8)
    img[..., 0] = np.broadcast_to(np.where(even, 200, 50), (h, w))
    img[..., 1] = np.broadcast_to(np.where(even, 50, 200), (h, w))
    img[..., 2] = 100
    img[..., 3] = 255
    return img


def minus_red(img: np.ndarray, delta: int) -> np.ndarray:
    """The perturbation of TestSSIMSimilar / TestMSSSIMSimilar (fennec_test.go:99-107,148-156):
    R -= delta wherever R > delta."""
    out = img.copy()
    r = out[..., 0]
    r[r > delta] -= np.uint8(delta)
    return out


def noise_image(w: int, h: int, seed: int, alpha: str = "opaque") -> np.ndarray:
    """Uniform uint8 RGB; alpha 'opaque' (255), 'random' or 'ramp'."""
    rng = np.random.Generator(np.random.PCG64(seed))
    img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    if alpha == "opaque":
        img[..., 3] = 255
    elif alpha == "ramp":
        img[..., 3] = make_test_image_with_alpha(w, h)[..., 3]
    return img


def perturb(img: np.ndarray, seed: int, amp: int = 6) -> np.ndarray:
    """B = clip(A + U{-amp..amp}) on RGB."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d = rng.integers(-amp, amp + 1, size=img.shape[:2] + (3,), dtype=np.int16)
    out = img.copy()
    out[..., :3] = np.clip(img[..., :3].astype(np.int16) + d, 0, 255).astype(np.uint8)
    return out


def gradient_noise_image(w: int, h: int, seed: int, sigma: float = 6.0) -> np.ndarray:
    """Smooth gradient + Gaussian noise: the 'synthetic photo' of configs 2/3."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = make_test_image(w, h)[..., :3].astype(np.float64)
    n = rng.normal(0.0, sigma, size=(h, w, 3))
    img = np.empty((h, w, 4), dtype=np.uint8)
    img[..., :3] = np.clip(np.rint(base + n), 0, 255).astype(np.uint8)
    img[..., 3] = 255
    return img


def flat_pm1_image(w: int, h: int, base: int, seed: int) -> np.ndarray:
    """Near-flat field base + U{-1,0,1}: the FP32 cancellation adversary."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d = rng.integers(-1, 2, size=(h, w, 3), dtype=np.int16)
    img = np.empty((h, w, 4), dtype=np.uint8)
    img[..., :3] = np.clip(base + d, 0, 255).astype(np.uint8)
    img[..., 3] = 255
    return img


def checker_flat_image(w: int, h: int, block: int, lo: int, hi: int, seed: int) -> np.ndarray:
    """Checkerboard of near-flat blocks at two far-apart levels (+-1 noise): stresses a
    per-tile centring constant because every tile sees both levels."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.arange(w)[None, :] // block
    y = np.arange(h)[:, None] // block
    level = np.where((x + y) % 2 == 0, lo, hi).astype(np.int16)
    d = rng.integers(-1, 2, size=(h, w, 3), dtype=np.int16)
    img = np.empty((h, w, 4), dtype=np.uint8)
    img[..., :3] = np.clip(level[..., None] + d, 0, 255).astype(np.uint8)
    img[..., 3] = 255
    return img


_SUBSAMPLE = {0: (1, 1), 1: (2, 1), 2: (2, 2), 3: (1, 2), 4: (4, 1), 5: (4, 2)}  # image.YCbCrSubsampleRatio -> (dx, dy)


def ycbcr_planes_from_nrgba(img: np.ndarray, ratio: int, seed: int = 0, amp: int = 0):
    """A deterministic stand-in for "encode then jpeg.Decode": JFIF forward transform (float, rounded), chroma
    averaged over each subsampling cell, optional +-amp noise on Y.  Only an input generator — the planes it
    returns are what the conversion under test consumes."""
    rgb = img[..., :3].astype(np.float64)
    h, w = img.shape[:2]
    y = 0.299 * rgb[..., 0] + 0.587 * rgb[..., 1] + 0.114 * rgb[..., 2]
    cb = 128.0 - 0.168736 * rgb[..., 0] - 0.331264 * rgb[..., 1] + 0.5 * rgb[..., 2]
    cr = 128.0 + 0.5 * rgb[..., 0] - 0.418688 * rgb[..., 1] - 0.081312 * rgb[..., 2]
    if amp:
        rng = np.random.Generator(np.random.PCG64(seed))
        y = y + rng.integers(-amp, amp + 1, size=y.shape)
    dx, dy = _SUBSAMPLE[ratio]
    cw, ch = (w + dx - 1) // dx, (h + dy - 1) // dy

    def sub(p):
        pad = np.pad(p, ((0, ch * dy - h), (0, cw * dx - w)), mode="edge")
        return pad.reshape(ch, dy, cw, dx).mean(axis=(1, 3))

    q = lambda p: np.clip(np.rint(p), 0, 255).astype(np.uint8)  # noqa: E731
    return q(y), q(sub(cb)), q(sub(cr))


def noise_planes(w: int, h: int, ratio: int, seed: int):
    """Uniform random Y/Cb/Cr planes (exercises the clamps of the colour transform)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    dx, dy = _SUBSAMPLE[ratio]
    cw, ch = (w + dx - 1) // dx, (h + dy - 1) // dy
    return (rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (ch, cw), dtype=np.uint8),
            rng.integers(0, 256, (ch, cw), dtype=np.uint8))
