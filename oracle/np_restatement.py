"""Second, independent restatement of the fennec hot path in NumPy float64 (TEST INFRASTRUCTURE ONLY).

Written from the Go sources separately from fennec_oracle.c and in a different style (vectorised
over pixels, sequential over taps) so that an agreement of the two pins the arithmetic as far as
it can be pinned without a Go toolchain (PARITY UNPINNED against the reference itself, see
fennec_oracle.h).  NumPy float64 elementwise ops are IEEE binary64 and never fused, and every
element sees the reference's operation order, so pixels must match the C oracle bit for bit and
scores to <= 1e-12.  Transcendentals go through Python's math module (glibc), like the C oracle.

Images: uint8 arrays (h, w, 4), NRGBA.
"""
from __future__ import annotations

import math

import numpy as np

C1 = 6.5025  # ssim.go:15 — exact constant arithmetic in Go, rounded once
C2 = 58.5225  # ssim.go:16


def go_round(x: np.ndarray) -> np.ndarray:
    """math.Round: nearest, halves away from zero (exact, no x+0.5 double rounding)."""
    x = np.asarray(x, dtype=np.float64)
    t = np.trunc(x)
    return t + np.sign(x) * (np.abs(x - t) >= 0.5)


def clampf(x: np.ndarray) -> np.ndarray:
    """convert.go:149-158"""
    return np.clip(go_round(x), 0, 255).astype(np.uint8)


def gaussian_kernel(size: int = 8, sigma: float = 1.5) -> np.ndarray:
    """ssim.go:223-241"""
    half = size // 2
    vals = []
    s = 0.0
    for y in range(-half, half):
        for x in range(-half, half):
            v = math.exp(-float(x * x + y * y) / (2 * sigma * sigma))
            vals.append(v)
            s += v
    return np.array([v / s for v in vals], dtype=np.float64)


def to_luminance(img: np.ndarray) -> np.ndarray:
    """ssim.go:207-220"""
    f = img.astype(np.float64)
    return (0.299 * f[..., 0] + 0.587 * f[..., 1]) + 0.114 * f[..., 2]


def _seq_sum(v: np.ndarray) -> float:
    v = np.ascontiguousarray(v, dtype=np.float64).ravel()
    if v.size == 0:
        return 0.0
    return float(np.cumsum(v)[-1])  # strictly left-to-right, unlike np.sum (pairwise)


def windowed_ssim(la: np.ndarray, lb: np.ndarray, procs: int = 8) -> float:
    """ssim.go:73-166 — all windows at once, taps in the reference's ki order."""
    h, w = la.shape
    k = gaussian_kernel(8, 1.5)
    oh, ow = h - 8, w - 8  # y in [4, h-4), x in [4, w-4)
    rows = h - 8 + 1
    procs = max(1, min(procs, rows))
    if oh <= 0 or ow <= 0:
        return 1.0
    mu_a = np.zeros((oh, ow))
    mu_b = np.zeros((oh, ow))
    ki = 0
    for wy in range(8):
        for wx in range(8):
            wt = k[ki]
            mu_a = mu_a + la[wy:wy + oh, wx:wx + ow] * wt
            mu_b = mu_b + lb[wy:wy + oh, wx:wx + ow] * wt
            ki += 1
    saa = np.zeros((oh, ow))
    sbb = np.zeros((oh, ow))
    sab = np.zeros((oh, ow))
    ki = 0
    for wy in range(8):
        for wx in range(8):
            wt = k[ki]
            da = la[wy:wy + oh, wx:wx + ow] - mu_a
            db = lb[wy:wy + oh, wx:wx + ow] - mu_b
            saa = saa + da * da * wt
            sbb = sbb + db * db * wt
            sab = sab + da * db * wt
            ki += 1
    num = (2 * mu_a * mu_b + C1) * (2 * sab + C2)
    den = (mu_a * mu_a + mu_b * mu_b + C1) * (saa + sbb + C2)
    smap = num / den
    rows_per = (rows + procs - 1) // procs
    total, count = 0.0, 0
    for p in range(procs):
        y0 = p * rows_per  # in window-row coordinates (y - 4)
        y1 = min(y0 + rows_per, oh)
        if y0 >= y1:
            continue
        total += _seq_sum(smap[y0:y1])
        count += (y1 - y0) * ow
    return 1.0 if count == 0 else total / float(count)


def pixel_ssim(a: np.ndarray, b: np.ndarray) -> float:
    """ssim.go:169-204"""
    h, w = a.shape[:2]
    n = float(w * h)
    if n == 0:
        return 1.0
    la = to_luminance(a).ravel()
    lb = to_luminance(b).ravel()
    mu_a = _seq_sum(la) / n
    mu_b = _seq_sum(lb) / n
    da = la - mu_a
    db = lb - mu_b
    saa = _seq_sum(da * da) / n
    sbb = _seq_sum(db * db) / n
    sab = _seq_sum(da * db) / n
    num = (2 * mu_a * mu_b + C1) * (2 * sab + C2)
    den = (mu_a * mu_a + mu_b * mu_b + C1) * (saa + sbb + C2)
    return num / den


def ssim(a: np.ndarray, b: np.ndarray, procs: int = 8) -> float:
    """ssim.go:24-43 (equal dims)"""
    h, w = a.shape[:2]
    if w < 8 or h < 8:
        return pixel_ssim(a, b)
    return windowed_ssim(to_luminance(a), to_luminance(b), procs)


def _box_edges(src: int, dst: int):
    """ssim.go:255-278 — int(float64(d)*ratio) truncation with the clamps."""
    ratio = float(src) / float(dst)
    lo, hi = [], []
    for d in range(dst):
        s0 = int(float(d) * ratio)
        s1 = int(float(d + 1) * ratio)
        if s1 > src:
            s1 = src
        if s0 >= s1:
            s0 = s1 - 1
        if s0 < 0:
            s0 = 0
        lo.append(s0)
        hi.append(s1)
    return np.array(lo), np.array(hi)


def box_downsample(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """ssim.go:244-309 via an exact integer summed-area table."""
    sh, sw = img.shape[:2]
    if sw <= 0 or sh <= 0 or dw <= 0 or dh <= 0:
        return np.zeros((0, 0, 4), dtype=np.uint8)
    x0, x1 = _box_edges(sw, dw)
    y0, y1 = _box_edges(sh, dh)
    sat = np.zeros((sh + 1, sw + 1, 4), dtype=np.int64)
    sat[1:, 1:] = np.cumsum(np.cumsum(img.astype(np.int64), axis=0), axis=1)
    s = (sat[y1][:, x1] - sat[y0][:, x1] - sat[y1][:, x0] + sat[y0][:, x0]).astype(np.float64)
    count = ((y1 - y0)[:, None] * (x1 - x0)[None, :]).astype(np.float64)
    out = np.zeros((dh, dw, 4), dtype=np.uint8)
    ok = count > 0
    inv = np.where(ok, 1.0 / np.where(ok, count, 1.0), 0.0)
    vals = clampf(s * inv[..., None])
    out[ok] = vals[ok]
    return out


def ssim_fast_dims(w: int, h: int):
    """ssim.go:52-56"""
    if w > 512 or h > 512:
        scale = 512.0 / max(float(w), float(h))
        nw = int(max(8.0, float(go_round(float(w) * scale))))
        nh = int(max(8.0, float(go_round(float(h) * scale))))
        return True, nw, nh
    return False, w, h


def ssim_fast(a: np.ndarray, b: np.ndarray, procs: int = 8) -> float:
    """ssim.go:48-70"""
    h, w = a.shape[:2]
    did, nw, nh = ssim_fast_dims(w, h)
    if did:
        a = box_downsample(a, nw, nh)
        b = box_downsample(b, nw, nh)
    return ssim(a, b, procs)


def msssim(a: np.ndarray, b: np.ndarray, procs: int = 8) -> float:
    """ssim.go:313-365 (equal dims)"""
    h, w = a.shape[:2]
    weights = [0.0448, 0.2856, 0.3001, 0.2363, 0.1333]
    levels = len(weights)
    for i in range(levels - 1):
        if min(w, h) < 8:
            weights = weights[: i + 1]
            s = 0.0
            for wt in weights:
                s += wt
            weights = [wt / s for wt in weights]
            break
        w //= 2
        h //= 2
    ca, cb = a.copy(), b.copy()
    result = 0.0
    for i, wt in enumerate(weights):
        s = ssim_fast(ca, cb, procs)
        result += wt * math.log(max(s, 1e-10))
        if i < len(weights) - 1:
            nw, nh = ca.shape[1] // 2, ca.shape[0] // 2
            if nw < 8 or nh < 8:
                break
            ca = box_downsample(ca, nw, nh)
            cb = box_downsample(cb, nw, nh)
    return math.exp(result)


def blur_kernel(sigma: float):
    """effects.go:153-165"""
    radius = int(math.ceil(sigma * 3))
    vals = []
    s = 0.0
    for i in range(2 * radius + 1):
        x = float(i - radius)
        v = math.exp(-(x * x) / (2 * sigma * sigma))
        vals.append(v)
        s += v
    return np.array([v / s for v in vals], dtype=np.float64), radius


def _blur_pass(planes: np.ndarray, kernel: np.ndarray, radius: int, axis: int) -> np.ndarray:
    n = planes.shape[axis]
    acc = np.zeros(planes.shape, dtype=np.float64)
    pos = np.arange(n)
    for k in range(2 * radius + 1):  # ascending k, clamp-to-edge (effects.go:172-184)
        idx = np.clip(pos + k - radius, 0, n - 1)
        acc = acc + np.take(planes, idx, axis=axis) * kernel[k]
    return acc


def gaussian_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    """effects.go:146-220 — uint8 intermediate between the passes, alpha from the source."""
    if sigma <= 0:
        return img
    kernel, radius = blur_kernel(sigma)
    rgb = img[..., :3].astype(np.float64)
    tmp = clampf(_blur_pass(rgb, kernel, radius, axis=1))
    out = np.empty_like(img)
    out[..., :3] = clampf(_blur_pass(tmp.astype(np.float64), kernel, radius, axis=0))
    out[..., 3] = img[..., 3]
    return out


def blur3x3(img: np.ndarray) -> np.ndarray:
    """effects.go:116-141"""
    h, w = img.shape[:2]
    out = img.copy()
    if h < 3 or w < 3:
        return out
    f = img[..., :3].astype(np.float64)
    wts = [(-1, -1, 1), (-1, 0, 2), (-1, 1, 1), (0, -1, 2), (0, 0, 4), (0, 1, 2), (1, -1, 1), (1, 0, 2), (1, 1, 1)]
    s = np.zeros((h - 2, w - 2, 3))
    for dy, dx, wt in wts:
        s = s + f[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx] * float(wt)
    out[1:h - 1, 1:w - 1, :3] = clampf(s / 16.0)
    return out


def sharpen(img: np.ndarray, strength: float) -> np.ndarray:
    """effects.go:10-45"""
    if strength <= 0:
        return img
    strength = min(strength, 1.0)
    h, w = img.shape[:2]
    if w < 3 or h < 3:
        return img
    blurred = blur3x3(img)
    amount = 1.0 + strength * 1.5
    orig = img[..., :3].astype(np.float64)
    bl = blurred[..., :3].astype(np.float64)
    out = np.empty_like(img)
    out[..., :3] = clampf(orig + amount * (orig - bl))
    out[..., 3] = img[..., 3]
    return out


def adaptive_sharpen(img: np.ndarray, strength: float) -> np.ndarray:
    """effects.go:49-112"""
    if strength <= 0:
        return img
    strength = min(strength, 1.0)
    h, w = img.shape[:2]
    if w < 3 or h < 3:
        return img
    blurred = blur3x3(img)
    amount = 1.0 + strength * 2.0
    lum = to_luminance(img)

    def L(dx, dy):
        return lum[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx]

    gx = -L(-1, -1) + L(1, -1) - 2 * L(-1, 0) + 2 * L(1, 0) - L(-1, 1) + L(1, 1)
    gy = -L(-1, -1) - 2 * L(0, -1) - L(1, -1) + L(-1, 1) + 2 * L(0, 1) + L(1, 1)
    edge = np.minimum(np.sqrt(gx * gx + gy * gy) / 400.0, 1.0)
    local = amount * edge
    out = img.copy()
    orig = img[1:h - 1, 1:w - 1, :3].astype(np.float64)
    bl = blurred[1:h - 1, 1:w - 1, :3].astype(np.float64)
    out[1:h - 1, 1:w - 1, :3] = clampf(orig + local[..., None] * (orig - bl))
    return out


def lanczos_kernel(x: float) -> float:
    """resize.go:57-69"""
    if x == 0:
        return 1.0
    if x < 0:
        x = -x
    if x >= 3.0:
        return 0.0
    xpi = x * math.pi
    return (3.0 * math.sin(xpi) * math.sin(xpi / 3.0)) / (xpi * xpi)


def lanczos_weights(dst_size: int, src_size: int):
    """resize.go:164-197 → list of (indices, weights) per destination index."""
    ratio = float(src_size) / float(dst_size)
    support = 3.0 * ratio if ratio > 1 else 3.0
    fscale = max(ratio, 1.0)
    table = []
    for d in range(dst_size):
        center = (float(d) + 0.5) * ratio - 0.5
        left = max(int(math.ceil(center - support)), 0)
        right = min(int(math.floor(center + support)), src_size - 1)
        idx, wts, wsum = [], [], 0.0
        for s in range(left, right + 1):
            wv = lanczos_kernel((float(s) - center) / fscale)
            if wv != 0:
                wsum += wv
                idx.append(s)
                wts.append(wv)
        if wsum != 0:
            wts = [v / wsum for v in wts]
        table.append((idx, wts))
    return table


def _resize_axis(img: np.ndarray, dst_size: int, axis: int) -> np.ndarray:
    """resize.go:77-118 (axis=1) / 121-161 (axis=0): premultiplied accumulate, a>0.5 gate."""
    src_size = img.shape[axis]
    table = lanczos_weights(dst_size, src_size)
    f = np.moveaxis(img.astype(np.float64), axis, 0)  # (src_size, other, 4)
    other = f.shape[1]
    out = np.zeros((dst_size, other, 4), dtype=np.uint8)
    for d, (idx, wts) in enumerate(table):
        r = np.zeros(other)
        g = np.zeros(other)
        b = np.zeros(other)
        a = np.zeros(other)
        for s, wv in zip(idx, wts):
            aw = f[s, :, 3] * wv
            r = r + f[s, :, 0] * aw
            g = g + f[s, :, 1] * aw
            b = b + f[s, :, 2] * aw
            a = a + aw
        ok = a > 0.5
        inv = 1.0 / np.where(ok, a, 1.0)
        px = np.stack([clampf(r * inv), clampf(g * inv), clampf(b * inv), clampf(a)], axis=-1)
        out[d][ok] = px[ok]
    return np.moveaxis(out, 0, axis)


def lanczos_resize(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """resize.go:37-53"""
    sh, sw = img.shape[:2]
    if sw <= 0 or sh <= 0 or dw <= 0 or dh <= 0:
        return np.zeros((0, 0, 4), dtype=np.uint8)
    if sw == dw and sh == dh:
        return img.copy()
    tmp = _resize_axis(img, dw, axis=1)
    return _resize_axis(tmp, dh, axis=0)


def smart_resize_dims(sw: int, sh: int, max_w: int, max_h: int):
    """resize.go:12-32"""
    if max_w <= 0:
        max_w = sw
    if max_h <= 0:
        max_h = sh
    if sw <= max_w and sh <= max_h:
        return True, sw, sh
    ratio = min(float(max_w) / float(sw), float(max_h) / float(sh))
    dw = int(max(1.0, float(go_round(float(sw) * ratio))))
    dh = int(max(1.0, float(go_round(float(sh) * ratio))))
    return False, dw, dh


# ---- §8(f1): convertToNRGBA (convert.go:34-64) on jpeg.Decode's output types ------------------------------
# Independent of the C oracle: chroma planes are upsampled by index arithmetic (Go's COffset), the colour
# transform is vectorised int32 (Go's color.YCbCr.RGBA(), image/color/ycbcr.go — stdlib, restated).

_SUB = {0: (1, 1), 1: (2, 1), 2: (2, 2), 3: (1, 2), 4: (4, 1), 5: (4, 2)}  # ratio -> (x divisor, y divisor)


def ycbcr_to_nrgba(y: np.ndarray, cb: np.ndarray, cr: np.ndarray, ratio: int) -> np.ndarray:
    h, w = y.shape
    dx, dy = _SUB[ratio]
    yi = (np.arange(h) // dy)[:, None]
    xi = (np.arange(w) // dx)[None, :]
    yy1 = y.astype(np.int32) * 0x10101
    cb1 = cb[yi, xi].astype(np.int32) - 128
    cr1 = cr[yi, xi].astype(np.int32) - 128

    def chan(v):
        inrange = (v.astype(np.uint32) & np.uint32(0xFF000000)) == 0
        c16 = np.where(inrange, v >> 8, np.where(v < 0, 0, 0xFFFF))      # ^(v>>31) & 0xffff
        return (c16 >> 8).astype(np.uint8)                              # convert.go:50-52

    out = np.empty((h, w, 4), np.uint8)
    out[..., 0] = chan(yy1 + 91881 * cr1)
    out[..., 1] = chan(yy1 - 22554 * cb1 - 46802 * cr1)
    out[..., 2] = chan(yy1 + 116130 * cb1)
    out[..., 3] = 255
    return out


def gray_to_nrgba(g: np.ndarray) -> np.ndarray:
    out = np.empty(g.shape + (4,), np.uint8)
    out[..., 0] = out[..., 1] = out[..., 2] = g
    out[..., 3] = 255
    return out


# ---- convertToNRGBA (convert.go:34-64) for the other decoder outputs ------------------------------------------
# Vectorised, independent of the C oracle: At().RGBA() per concrete type (image/color), then convert.go:42-60.
FMT_RGBA, FMT_RGBA64, FMT_NRGBA64, FMT_GRAY16, FMT_CMYK, FMT_PALETTED = 1, 2, 3, 4, 5, 6


def _be16(a: np.ndarray) -> np.ndarray:
    """(..., 2k) big-endian bytes -> (..., k) uint64 values."""
    a = a.astype(np.uint64)
    return (a[..., 0::2] << np.uint64(8)) | a[..., 1::2]


def convert_to_nrgba(fmt: int, pix: np.ndarray, pal16: np.ndarray = None) -> np.ndarray:
    """pix: (h, w, 4) uint8 for RGBA / CMYK, (h, w, 8) for the 64-bit types, (h, w, 2) for Gray16, (h, w) indices
    for Paletted (pal16: (n, 4) uint16 = Palette[i].RGBA())."""
    u = np.uint64
    if fmt == FMT_RGBA:
        v = pix.astype(u)
        rgba = v | (v << u(8))
    elif fmt == FMT_RGBA64:
        rgba = _be16(pix)
    elif fmt == FMT_NRGBA64:
        v = _be16(pix)
        a = v[..., 3:4]
        rgba = np.concatenate([v[..., :3] * a // u(0xFFFF), a], -1)
    elif fmt == FMT_GRAY16:
        y = _be16(pix)
        rgba = np.concatenate([y, y, y, np.full_like(y, 0xFFFF)], -1)
    elif fmt == FMT_CMYK:
        v = pix.astype(u)
        wk = u(0xFFFF) - v[..., 3:4] * u(0x101)
        rgb = (u(0xFFFF) - v[..., :3] * u(0x101)) * wk // u(0xFFFF)
        rgba = np.concatenate([rgb, np.full_like(wk, 0xFFFF)], -1)
    elif fmt == FMT_PALETTED:
        assert pix.max(initial=0) < len(pal16), "Go panics: index out of range"
        rgba = pal16.astype(u)[pix]
    else:
        raise ValueError(fmt)
    a = rgba[..., 3:4]
    safe = np.where(a == 0, u(1), a)
    mid = (((rgba[..., :3] * u(0xFFFF)) // safe) >> u(8)) & u(0xFF)         # uint8() keeps the low byte
    rgb = np.where(a == 0xFFFF, rgba[..., :3] >> u(8), mid)
    out = np.concatenate([rgb, a >> u(8)], -1)
    out = np.where(a == 0, u(0), out)
    return out.astype(np.uint8)


# ---- §8(f2): Analyze (analyze.go:26-176) --------------------------------------------------------------------
# Independent restatement: vectorised per-pixel float64 luminance, np.cumsum for the SEQUENTIAL sums (cumsum adds in
# order, unlike np.sum's pairwise reduction), a dict-free distinct count for the capped colour set.

def _lum(img: np.ndarray) -> np.ndarray:
    p = img.astype(np.float64)
    return (0.299 * p[..., 0] + 0.587 * p[..., 1]) + 0.114 * p[..., 2]


def _go_log2(x: float) -> float:
    frac, e = math.frexp(x)
    if frac == 0.5:
        return float(e - 1)
    return math.log(frac) * (1.0 / math.log(2.0)) + float(e)


def analyze(img: np.ndarray) -> dict:
    h, w = img.shape[:2]
    st = dict(width=w, height=h, has_alpha=0, is_grayscale=0, unique_colors=0, entropy=0.0, edge_density=0.0,
              mean_brightness=0.0, contrast=0.0)
    if w == 0 or h == 0:
        return st
    lum = _lum(img)
    n = float(w * h)
    st["mean_brightness"] = float(np.cumsum(lum.ravel())[-1]) / n
    hist = np.bincount((lum + 0.5).astype(np.int64).ravel(), minlength=256).astype(np.float64)
    st["histogram"] = hist
    st["has_alpha"] = int((img[..., 3] < 255).any())
    st["is_grayscale"] = int(((img[..., 0] == img[..., 1]) & (img[..., 1] == img[..., 2])).all())
    step = (w * h) // 50000 if w * h > 50000 else 1
    flat = img.reshape(-1, 4)[::step].astype(np.uint32)
    keys = flat[:, 0] << 24 | flat[:, 1] << 16 | flat[:, 2] << 8 | flat[:, 3]
    # the map stops growing at 1024 entries: the count is min(distinct among the sampled keys, 1024)
    st["unique_colors"] = int(min(len(np.unique(keys)), 1024))
    sy, sx = int(max(1, math.ceil(h / 100))), int(max(1, math.ceil(w / 100)))
    d = lum[::sy, ::sx] - st["mean_brightness"]
    dd = (d * d).ravel()
    st["contrast"] = math.sqrt(float(np.cumsum(dd)[-1]) / float(dd.size))
    ent = 0.0
    for c in hist:
        if c > 0:
            p = c / n
            ent -= p * _go_log2(p)
    st["entropy"] = ent
    if w >= 3 and h >= 3:
        ex, ey = int(max(1, w / 200)), int(max(1, h / 200))
        ys, xs = np.arange(1, h - 1, ey), np.arange(1, w - 1, ex)
        L = lambda dx, dy: lum[np.ix_(ys + dy, xs + dx)]  # noqa: E731
        gx = ((((L(1, -1) - L(-1, -1)) + 2 * L(1, 0)) - 2 * L(-1, 0)) + L(1, 1)) - L(-1, 1)
        gy = ((((L(-1, 1) - L(-1, -1)) + 2 * L(0, 1)) - 2 * L(0, -1)) + L(1, 1)) - L(1, -1)
        mag = np.sqrt(gx * gx + gy * gy)
        st["edge_density"] = float((mag > 30.0).sum()) / float(mag.size)
    return st


# ---- §8(f4): ApplyOrientation (exif.go:176-203) with array operations instead of the reference's index loops ----

def apply_orientation(img: np.ndarray, orient: int) -> np.ndarray:
    rot90cw = lambda a: np.rot90(a, k=-1)      # noqa: E731  convert.go:186-198
    rot270cw = lambda a: np.rot90(a, k=1)      # noqa: E731  convert.go:216-226
    fliph = lambda a: a[:, ::-1]               # noqa: E731  convert.go:229-241
    ops = {2: fliph, 3: lambda a: a[::-1, ::-1], 4: lambda a: a[::-1], 5: lambda a: fliph(rot270cw(a)), 6: rot90cw,
           7: lambda a: fliph(rot90cw(a)), 8: rot270cw}
    if orient not in ops:
        return img
    return np.ascontiguousarray(ops[orient](img))


# ---- §8(f3): applyPalette + palettedToNRGBA (targetsize.go:479-545): distance matrix + argmin (first minimum) ----

def apply_palette(img: np.ndarray, palette: np.ndarray):
    pal = np.asarray(palette, dtype=np.int64)[:, :3]
    h, w = img.shape[:2]
    idx = np.empty((h, w), np.uint8)
    for y0 in range(0, h, 64):                      # row blocks keep the (rows, w, ncolors) tensor small
        px = img[y0:y0 + 64, :, :3].astype(np.int64)
        d = ((px[:, :, None, :] - pal[None, None, :, :]) ** 2).sum(-1)
        idx[y0:y0 + 64] = np.argmin(d, axis=-1).astype(np.uint8)
    return idx, np.asarray(palette, dtype=np.uint8)[idx]
