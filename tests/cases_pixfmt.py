"""Inputs for convertToNRGBA on the decoder output types other than YCbCr / Gray (tests/test_pixfmt.py)."""
import numpy as np

FMT_RGBA, FMT_RGBA64, FMT_NRGBA64, FMT_GRAY16, FMT_CMYK, FMT_PALETTED = 1, 2, 3, 4, 5, 6
BPP = {1: 4, 2: 8, 3: 8, 4: 2, 5: 4, 6: 1}
NAMES = {1: "RGBA", 2: "RGBA64", 3: "NRGBA64", 4: "Gray16", 5: "CMYK", 6: "Paletted"}


def _be(v16: np.ndarray) -> np.ndarray:
    """(..., k) uint16 values -> (..., 2k) big-endian bytes (Go's 64-bit / Gray16 Pix layout)."""
    out = np.empty(v16.shape[:-1] + (v16.shape[-1] * 2,), np.uint8)
    out[..., 0::2] = v16 >> 8
    out[..., 1::2] = v16 & 0xFF
    return out


def make(fmt: int, w: int, h: int, seed: int, kind: str = "valid"):
    """kind 'valid': premultiplied types respect c <= a (what decoders and draw ops produce), with plenty of a == 0,
    a == max and small a; 'wild': arbitrary bytes (c > a wraps through Go's uint8() truncation)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pal = None
    if fmt in (FMT_RGBA, FMT_CMYK):
        pix = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        if fmt == FMT_RGBA:
            a = rng.choice(np.array([0, 1, 2, 3, 127, 128, 254, 255], np.uint8), (h, w))
            a = np.where(rng.random((h, w)) < 0.5, a, rng.integers(0, 256, (h, w), dtype=np.uint8)).astype(np.uint8)
            pix[..., 3] = a
            if kind == "valid":
                pix[..., :3] = (pix[..., :3].astype(np.uint16) * a[..., None] // 255).astype(np.uint8)
    elif fmt in (FMT_RGBA64, FMT_NRGBA64):
        v = rng.integers(0, 65536, (h, w, 4), dtype=np.uint16)
        a = rng.choice(np.array([0, 1, 255, 256, 257, 0x7FFF, 0x8000, 0xFFFE, 0xFFFF], np.uint16), (h, w))
        a = np.where(rng.random((h, w)) < 0.5, a, rng.integers(0, 65536, (h, w), dtype=np.uint16)).astype(np.uint16)
        v[..., 3] = a
        if fmt == FMT_RGBA64 and kind == "valid":
            v[..., :3] = (v[..., :3].astype(np.uint64) * a[..., None] // 0xFFFF).astype(np.uint16)
        pix = _be(v)
    elif fmt == FMT_GRAY16:
        pix = _be(rng.integers(0, 65536, (h, w, 1), dtype=np.uint16))
    else:
        n = 256 if seed % 2 == 0 else 37
        pix = rng.integers(0, n, (h, w), dtype=np.uint8)
        pal = rng.integers(0, 65536, (n, 4), dtype=np.uint16)
        pal[::3, 3] = 0xFFFF                      # png without tRNS: opaque color.RGBA entries
        pal[1::7, 3] = 0
        if kind == "valid":
            pal[:, :3] = (pal[:, :3].astype(np.uint64) * pal[:, 3:4] // 0xFFFF).astype(np.uint16)
    return np.ascontiguousarray(pix), pal


def exhaustive_rgba():
    """Every (colour, alpha) byte pair of *image.RGBA: 256 x 256 pixels, R = G = B = column, A = row."""
    c, a = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8))
    return np.ascontiguousarray(np.stack([c, c, c, a], -1))
