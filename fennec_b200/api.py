"""Host-side mirror of the reference's Go API for the hot path, over the C ABI.

Same names, argument meaning and guard behaviour as the Go functions (file:line cited per
function), so the parity tests read like fennec_test.go.  Images are numpy uint8 arrays of shape
(h, w, 4) — NRGBA, `arr.strides[0]` plays image.NRGBA.Stride.  Where the reference returns its
input pointer unchanged, the SAME array object is returned.

Every function goes through libfennec_b200.so; there is no NumPy/CPU compute path here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import FB_IDENTITY, FbWeights, check, dp, ip, u8p


def _img(a: np.ndarray):
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 4):
        raise TypeError("expected a uint8 array of shape (h, w, 4) (NRGBA)")
    if a.size and (a.strides[2] != 1 or a.strides[1] != 4):
        raise ValueError("pixels must be interleaved NRGBA bytes")
    h, w = a.shape[:2]
    stride = int(a.strides[0]) if h > 1 else w * 4
    return a.ctypes.data_as(u8p), stride, w, h


def _new(h: int, w: int) -> np.ndarray:
    return np.zeros((h, w, 4), dtype=np.uint8)  # image.NewNRGBA zero-fills


def _empty() -> np.ndarray:
    return np.zeros((0, 0, 4), dtype=np.uint8)  # image.NewNRGBA(image.Rect(0,0,0,0))


def device_count() -> int:
    return _lib.load().fb_device_count()


def version() -> str:
    return _lib.load().fb_version().decode()


def set_device(device: int) -> None:
    check(_lib.load().fb_set_device(device))


# ---- resize.go ---------------------------------------------------------------------------------

def lanczos_weights(dst_size: int, src_size: int):
    """precomputeWeights (resize.go:164-197) as CSR arrays (start, index, weight)."""
    L = _lib.load()
    cap = L.fb_lanczos_weights_cap(dst_size, src_size)
    start = np.zeros(dst_size + 1, dtype=np.int32)
    index = np.zeros(cap + 1, dtype=np.int32)
    weight = np.zeros(cap + 1, dtype=np.float64)
    n = check(L.fb_build_lanczos_weights(dst_size, src_size, start.ctypes.data_as(ip), index.ctypes.data_as(ip),
                                         weight.ctypes.data_as(dp)))
    return start, index[:n].copy(), weight[:n].copy()


def _weights_struct(tab) -> Tuple[FbWeights, tuple]:
    start, index, weight = (np.ascontiguousarray(tab[0], dtype=np.int32), np.ascontiguousarray(tab[1], dtype=np.int32),
                            np.ascontiguousarray(tab[2], dtype=np.float64))
    w = FbWeights(len(start) - 1, start.ctypes.data_as(ip), index.ctypes.data_as(ip), weight.ctypes.data_as(dp))
    return w, (start, index, weight)


def lanczos_resize(img: np.ndarray, dst_w: int, dst_h: int, weights_x=None, weights_y=None) -> np.ndarray:
    """lanczosResize (resize.go:37-53). weights_* optionally carry caller-built CSR tables."""
    ps, ss, sw, sh = _img(img)
    if sw <= 0 or sh <= 0 or dst_w <= 0 or dst_h <= 0:
        return _empty()  # resize.go:41-43
    dst = _new(dst_h, dst_w)
    pd, sd, _, _ = _img(dst)
    wx = wy = None
    keep = []
    if weights_x is not None:
        wx, k = _weights_struct(weights_x)
        keep.append(k)
    if weights_y is not None:
        wy, k = _weights_struct(weights_y)
        keep.append(k)
    check(_lib.load().fb_lanczos_resize(ps, ss, sw, sh, pd, sd, dst_w, dst_h,
                                        C.byref(wx) if wx is not None else None,
                                        C.byref(wy) if wy is not None else None))
    return dst


def smart_resize(img: np.ndarray, max_w: int, max_h: int) -> np.ndarray:
    """smartResize (resize.go:12-32): returns `img` itself when it already fits."""
    _, _, sw, sh = _img(img)
    dw, dh = C.c_int(), C.c_int()
    if _lib.load().fb_smart_resize_dims(sw, sh, max_w, max_h, C.byref(dw), C.byref(dh)) == 1:
        return img
    return lanczos_resize(img, dw.value, dh.value)


# ---- ssim.go ------------------------------------------------------------------------------------

def _score(fn, a: np.ndarray, b: np.ndarray) -> float:
    pa, sa, w, h = _img(a)
    pb, sb, wb, hb = _img(b)
    if (w, h) != (wb, hb):
        raise ValueError("images must have equal dimensions")
    out = C.c_double()
    check(fn(pa, sa, pb, sb, w, h, C.byref(out)))
    return out.value


def SSIM(img1: np.ndarray, img2: np.ndarray) -> float:
    """fennec.SSIM (ssim.go:24-43): resizes img2 to img1's dims if they differ (ssim.go:31-33)."""
    h, w = img1.shape[:2]
    if img2.shape[:2] != (h, w):
        img2 = lanczos_resize(img2, w, h)
    return _score(_lib.load().fb_ssim, img1, img2)


def SSIMFast(img1: np.ndarray, img2: np.ndarray) -> float:
    """fennec.SSIMFast (ssim.go:48-70). Like the reference, equal dims are the caller's business."""
    return _score(_lib.load().fb_ssim_fast, img1, img2)


def MSSSIM(img1: np.ndarray, img2: np.ndarray) -> float:
    """fennec.MSSSIM (ssim.go:313-365)."""
    h, w = img1.shape[:2]
    if img2.shape[:2] != (h, w):
        img2 = lanczos_resize(img2, w, h)
    return _score(_lib.load().fb_msssim, img1, img2)


def pixel_ssim(a: np.ndarray, b: np.ndarray) -> float:
    """pixelSSIM (ssim.go:169-204)."""
    return _score(_lib.load().fb_pixel_ssim, a, b)


def box_downsample(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """boxDownsample (ssim.go:244-309)."""
    ps, ss, sw, sh = _img(img)
    if sw <= 0 or sh <= 0 or dst_w <= 0 or dst_h <= 0:
        return _empty()
    dst = _new(dst_h, dst_w)
    pd, sd, _, _ = _img(dst)
    check(_lib.load().fb_box_downsample(ps, ss, sw, sh, pd, sd, dst_w, dst_h))
    return dst


def ssim_fast_dims(w: int, h: int):
    nw, nh = C.c_int(), C.c_int()
    did = _lib.load().fb_ssim_fast_dims(w, h, C.byref(nw), C.byref(nh))
    return bool(did), nw.value, nh.value


# ---- effects.go ---------------------------------------------------------------------------------

def blur_kernel(sigma: float):
    """The 1-D kernel of effects.go:153-165 → (weights, radius)."""
    if not sigma > 0:
        raise ValueError("sigma must be > 0")
    L = _lib.load()
    need = -L.fb_blur_kernel(float(sigma), None, 0)
    k = np.zeros(need, dtype=np.float64)
    radius = check(L.fb_blur_kernel(float(sigma), k.ctypes.data_as(dp), need))
    return k, radius


def GaussianBlur(img: np.ndarray, sigma: float, kernel: Optional[np.ndarray] = None) -> np.ndarray:
    """fennec.GaussianBlur (effects.go:146-220). sigma <= 0 → the same array (effects.go:147-149).
    `kernel` lets the caller supply the weight table it built with its own libm."""
    if sigma <= 0:
        return img
    ps, ss, w, h = _img(img)
    dst = _new(h, w)
    pd, sd, _, _ = _img(dst)
    if kernel is None:
        check(_lib.load().fb_gaussian_blur_sigma(ps, ss, w, h, float(sigma), pd, sd))
    else:
        k = np.ascontiguousarray(kernel, dtype=np.float64)
        check(_lib.load().fb_gaussian_blur(ps, ss, w, h, k.ctypes.data_as(dp), (len(k) - 1) // 2, pd, sd))
    return dst


def blur3x3(img: np.ndarray) -> np.ndarray:
    """gaussianBlur3x3 (effects.go:116-141)."""
    ps, ss, w, h = _img(img)
    dst = _new(h, w)
    pd, sd, _, _ = _img(dst)
    check(_lib.load().fb_blur3x3(ps, ss, w, h, pd, sd))
    return dst


def _fx(fn, img: np.ndarray, strength: float) -> np.ndarray:
    ps, ss, w, h = _img(img)
    dst = _new(h, w)
    pd, sd, _, _ = _img(dst)
    if check(fn(ps, ss, w, h, float(strength), pd, sd)) == FB_IDENTITY:
        return img  # same pointer (effects.go:11-22 / 50-61)
    return dst


def Sharpen(img: np.ndarray, strength: float) -> np.ndarray:
    """fennec.Sharpen (effects.go:10-45)."""
    return _fx(_lib.load().fb_sharpen, img, strength)


def AdaptiveSharpen(img: np.ndarray, strength: float) -> np.ndarray:
    """fennec.AdaptiveSharpen (effects.go:49-90)."""
    return _fx(_lib.load().fb_adaptive_sharpen, img, strength)


# ---- convert.go:34-64 and the quality search's reference session (compress.go:45-74) ------

# image.YCbCrSubsampleRatio constants, in Go's order
YCBCR_444, YCBCR_422, YCBCR_420, YCBCR_440, YCBCR_411, YCBCR_410 = range(6)
_SUBSAMPLE = {0: (1, 1), 1: (2, 1), 2: (2, 2), 3: (1, 2), 4: (4, 1), 5: (4, 2)}


def chroma_dims(w: int, h: int, ratio: int) -> Tuple[int, int]:
    """Plane size image.NewYCbCr allocates for a (0,0)-(w,h) rectangle."""
    dx, dy = _SUBSAMPLE[ratio]
    return (w + dx - 1) // dx, (h + dy - 1) // dy


def _plane(a: np.ndarray):
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.ndim == 2 and (a.size == 0 or a.strides[1] == 1)):
        raise TypeError("expected a 2-D uint8 plane with contiguous rows")
    return a.ctypes.data_as(u8p), (int(a.strides[0]) if a.shape[0] > 1 else a.shape[1])


def ycbcr_to_nrgba(y: np.ndarray, cb: np.ndarray, cr: np.ndarray, ratio: int) -> np.ndarray:
    """convertToNRGBA (convert.go:34-64) of a decoded *image.YCbCr (Y, Cb, Cr planes + subsample ratio)."""
    h, w = y.shape
    if cb.shape != cr.shape or cb.strides != cr.strides:
        raise ValueError("Cb and Cr must share shape and stride (image.YCbCr.CStride)")
    dst = _new(h, w)
    py, sy = _plane(y)
    pcb, sc = _plane(cb)
    pcr, _ = _plane(cr)
    pd, sd, _, _ = _img(dst)
    check(_lib.load().fb_ycbcr_to_nrgba(py, sy, pcb, pcr, sc, w, h, ratio, pd, sd))
    return dst


def gray_to_nrgba(g: np.ndarray) -> np.ndarray:
    """convertToNRGBA (convert.go:34-64) of a decoded *image.Gray."""
    h, w = g.shape
    dst = _new(h, w)
    pg, sg = _plane(g)
    pd, sd, _, _ = _img(dst)
    check(_lib.load().fb_gray_to_nrgba(pg, sg, w, h, pd, sd))
    return dst


FMT_RGBA, FMT_RGBA64, FMT_NRGBA64, FMT_GRAY16, FMT_CMYK, FMT_PALETTED = 1, 2, 3, 4, 5, 6
_FMT_BPP = {FMT_RGBA: 4, FMT_RGBA64: 8, FMT_NRGBA64: 8, FMT_GRAY16: 2, FMT_CMYK: 4, FMT_PALETTED: 1}


def convert_to_nrgba(fmt: int, pix: np.ndarray, palette16: Optional[np.ndarray] = None) -> np.ndarray:
    """convertToNRGBA (convert.go:34-64) of a decoded *image.RGBA / RGBA64 / NRGBA64 / Gray16 / CMYK / Paletted:
    `pix` is the Go image's Pix as (h, w, bytes-per-pixel) uint8 — (h, w) for Paletted, whose `palette16` holds
    Palette[i].RGBA() as (n, 4) uint16.  Rows may be strided; pixels are packed within a row."""
    if fmt not in _FMT_BPP:
        raise ValueError(f"unknown pixel format {fmt}")
    bpp = _FMT_BPP[fmt]
    if pix.ndim == 2:
        pix = pix[..., None]
    if pix.dtype != np.uint8 or pix.ndim != 3 or pix.shape[2] != bpp:
        raise TypeError(f"expected uint8 pixels of shape (h, w, {bpp})")
    h, w = pix.shape[:2]
    if (bpp > 1 and pix.strides[2] != 1) or (w > 1 and pix.strides[1] != bpp):
        raise ValueError("pixels must be packed within a row")
    dst = _new(h, w)
    pd, sd, _, _ = _img(dst)
    pp, n = None, 0
    if palette16 is not None:
        palette16 = np.ascontiguousarray(palette16, dtype=np.uint16)
        if palette16.ndim != 2 or palette16.shape[1] != 4:
            raise TypeError("palette16 must have shape (n, 4)")
        pp, n = palette16.ctypes.data_as(C.POINTER(C.c_uint16)), len(palette16)
    stride = int(pix.strides[0]) if h > 1 else w * bpp
    check(_lib.load().fb_convert_to_nrgba(fmt, pix.ctypes.data_as(_lib.u8p), stride, w, h, pp, n, pd, sd))
    return dst


class SSIMReference:
    """The `src` side of compress.go:45-74's search, kept on the device: SSIMFast(src, candidate) per iteration
    with only the candidate crossing PCIe (as YCbCr planes or NRGBA)."""

    def __init__(self, src: np.ndarray):
        p, stride, w, h = _img(src)
        self._h = C.c_void_p()
        self.w, self.h = w, h
        check(_lib.load().fb_ssim_ref_create(p, stride, w, h, C.byref(self._h)))

    def score_ycbcr(self, y: np.ndarray, cb: np.ndarray, cr: np.ndarray, ratio: int) -> float:
        if y.shape != (self.h, self.w):
            raise ValueError("candidate dims differ from the reference image")
        py, sy = _plane(y)
        pcb, sc = _plane(cb)
        pcr, _ = _plane(cr)
        out = C.c_double()
        check(_lib.load().fb_ssim_ref_score_ycbcr(self._h, py, sy, pcb, pcr, sc, ratio, C.byref(out)))
        return out.value

    def score_nrgba(self, img: np.ndarray) -> float:
        p, stride, w, h = _img(img)
        if (w, h) != (self.w, self.h):
            raise ValueError("candidate dims differ from the reference image")
        out = C.c_double()
        check(_lib.load().fb_ssim_ref_score_nrgba(self._h, p, stride, C.byref(out)))
        return out.value

    def close(self) -> None:
        if self._h:
            _lib.load().fb_ssim_ref_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- analyze.go:26-176 -----------------------------------------------------------------------------

FORMAT_JPEG, FORMAT_PNG = 1, 2                                  # types.go:36-42
QUALITY_BALANCED, QUALITY_HIGH, QUALITY_AGGRESSIVE = 0, 3, 4    # types.go:59-70


def _stats_dict(st: "_lib.FbImageStats") -> dict:
    return {k: getattr(st, k) for k, _ in _lib.FbImageStats._fields_}


def Analyze(img: np.ndarray) -> dict:
    """fennec.Analyze (analyze.go:26-113) → the ImageStats fields as a dict (snake_case keys)."""
    st = _lib.FbImageStats()
    if img.size == 0:
        check(_lib.load().fb_analyze(None, 0, img.shape[1] if img.ndim == 3 else 0, img.shape[0] if img.ndim == 3 else 0, C.byref(st)))
    else:
        p, stride, w, h = _img(img)
        check(_lib.load().fb_analyze(p, stride, w, h, C.byref(st)))
    return _stats_dict(st)


# ---- exif.go:176-203 ---------------------------------------------------------------------------------

ORIENT_NORMAL, ORIENT_FLIP_H, ORIENT_ROTATE_180, ORIENT_FLIP_V = 1, 2, 3, 4          # exif.go:12-21
ORIENT_TRANSPOSE, ORIENT_ROTATE_90CW, ORIENT_TRANSVERSE, ORIENT_ROTATE_270CW = 5, 6, 7, 8


def ApplyOrientation(img: np.ndarray, orient: int) -> np.ndarray:
    """fennec.ApplyOrientation (exif.go:176-203); orientations 1, 0 and unknown values return `img` itself."""
    p, stride, w, h = _img(img)
    dw, dh = C.c_int(), C.c_int()
    if check(_lib.load().fb_orientation_dims(orient, w, h, C.byref(dw), C.byref(dh))) == FB_IDENTITY:
        return img
    dst = _new(dh.value, dw.value)
    pd, sd, _, _ = _img(dst)
    check(_lib.load().fb_apply_orientation(p, stride, w, h, orient, pd, sd if dst.size else dw.value * 4))
    return dst


# ---- targetsize.go:479-545 ---------------------------------------------------------------------------

def apply_palette(src: np.ndarray, palette: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """applyPalette + palettedToNRGBA (targetsize.go:479-545): (indices (h, w) uint8, reconstruction (h, w, 4)).
    `palette` is (ncolors, 4) uint8 NRGBA with alpha 255 — what medianCut returns."""
    pal = np.ascontiguousarray(palette, dtype=np.uint8)
    if pal.ndim != 2 or pal.shape[1] != 4:
        raise TypeError("palette must have shape (ncolors, 4)")
    p, stride, w, h = _img(src)
    idx = np.zeros((h, w), dtype=np.uint8)
    dst = _new(h, w)
    pd, sd, _, _ = _img(dst)
    check(_lib.load().fb_apply_palette(p, stride, w, h, pal.ctypes.data_as(u8p), pal.shape[0], idx.ctypes.data_as(u8p), w, pd, sd))
    return idx, dst


# ---- batch.go: CompressBatch's worker pool behind one call (host buffers, every initialised GPU) ---------------

class _BatchOpts:
    """Builds struct fb_batch_opts and keeps its ctypes objects alive for the duration of the call."""

    def __init__(self, workers_per_device=0, cancel=None, on_item=None):
        self.cancel = cancel if cancel is not None else C.c_int(0)
        self._cb = _lib.PROGRESS_FN(lambda done, total, _u: on_item(done, total)) if on_item else _lib.PROGRESS_FN()
        self.struct = _lib.FbBatchOpts(int(workers_per_device), C.pointer(self.cancel), self._cb, None)


def _statuses(n):
    return (C.c_int * max(n, 1))()


def score_batch(op: str, pairs, workers_per_device: int = 0, cancel=None, on_item=None):
    """fb_score_batch_host: fennec.SSIM / SSIMFast / MSSSIM for a list of (img1, img2) host pairs, sharded over every
    initialised GPU from this one process (batch.go:58-128).  Returns (scores, statuses) in input order; items whose
    status is negative (failed or cancelled) carry NaN."""
    code = {"ssim": _lib.FB_OP_SSIM, "ssim_fast": _lib.FB_OP_SSIM_FAST, "msssim": _lib.FB_OP_MSSSIM}[op]
    n = len(pairs)
    arr = (_lib.FbPair * max(n, 1))()
    for i, (a, b) in enumerate(pairs):
        pa, sa, w, h = _img(a)
        pb, sb, wb, hb = _img(b)
        if (w, h) != (wb, hb):
            raise ValueError(f"pair {i}: images must have equal dimensions")
        arr[i] = _lib.FbPair(C.cast(pa, C.c_void_p), sa, C.cast(pb, C.c_void_p), sb, w, h)
    scores = np.full(max(n, 1), np.nan, dtype=np.float64)
    st = _statuses(n)
    opts = _BatchOpts(workers_per_device, cancel, on_item)
    check(_lib.load().fb_score_batch_host(code, arr, n, scores.ctypes.data_as(dp), st, C.byref(opts.struct)))
    st = list(st)[:n]
    scores = scores[:n]
    scores[[i for i, s_ in enumerate(st) if s_ < 0]] = np.nan
    return scores, st


def lanczos_resize_batch(imgs, dst_w: int, dst_h: int, workers_per_device: int = 0, cancel=None, on_item=None):
    """fb_lanczos_resize_batch_host: lanczosResize (resize.go:37-53) of every image of a list → (outputs, statuses)."""
    n = len(imgs)
    arr = (_lib.FbResizeItem * max(n, 1))()
    outs = []
    for i, im in enumerate(imgs):
        ps, ss, sw, sh = _img(im)
        dst = _new(max(dst_h, 0), max(dst_w, 0))
        pd, sd, _, _ = _img(dst)
        outs.append(dst)
        arr[i] = _lib.FbResizeItem(C.cast(ps, C.c_void_p), ss, sw, sh, C.cast(pd, C.c_void_p), sd, dst_w, dst_h)
    st = _statuses(n)
    opts = _BatchOpts(workers_per_device, cancel, on_item)
    check(_lib.load().fb_lanczos_resize_batch_host(arr, n, st, C.byref(opts.struct)))
    st = list(st)[:n]
    return [(_empty() if s_ == FB_IDENTITY else o) for o, s_ in zip(outs, st)], st


def effect_batch(effect: str, param: float, imgs, workers_per_device: int = 0, cancel=None, on_item=None):
    """fb_effect_batch_host: GaussianBlur(sigma) / Sharpen(strength) / AdaptiveSharpen(strength) of every image of a
    list; where the reference returns its input pointer (effects.go:11-22,147-149) the SAME array object comes back."""
    code = {"gaussian_blur": _lib.FB_FX_GAUSSIAN_BLUR, "sharpen": _lib.FB_FX_SHARPEN,
            "adaptive_sharpen": _lib.FB_FX_ADAPTIVE_SHARPEN}[effect]
    n = len(imgs)
    arr = (_lib.FbEffectItem * max(n, 1))()
    outs = []
    for i, im in enumerate(imgs):
        ps, ss, w, h = _img(im)
        dst = _new(h, w)
        pd, sd, _, _ = _img(dst)
        outs.append(dst)
        arr[i] = _lib.FbEffectItem(C.cast(ps, C.c_void_p), ss, C.cast(pd, C.c_void_p), sd, w, h)
    st = _statuses(n)
    opts = _BatchOpts(workers_per_device, cancel, on_item)
    check(_lib.load().fb_effect_batch_host(code, float(param), arr, n, st, C.byref(opts.struct)))
    st = list(st)[:n]
    return [(im if s_ == FB_IDENTITY else o) for im, o, s_ in zip(imgs, outs, st)], st


def init(devices=None) -> int:
    """fb_init: select the GPUs this process uses (None = all visible); returns the device count."""
    if devices is None:
        return check(_lib.load().fb_init(None, 0))
    arr = (C.c_int * len(devices))(*devices)
    return check(_lib.load().fb_init(arr, len(devices)))


def shutdown() -> None:
    _lib.load().fb_shutdown()


def pinned_empty(shape) -> np.ndarray:
    """A uint8 array in page-locked host memory (fb_alloc_pinned): uploads from it are DMA'ed without staging.  The
    memory is released when the array (and every view of it) is garbage-collected."""
    import weakref
    nbytes = int(np.prod(shape))
    L = _lib.load()
    ptr = L.fb_alloc_pinned(nbytes)
    if not ptr:
        check(_lib.FB_E_OOM)
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.uint8, count=nbytes).reshape(shape)
    weakref.finalize(buf, L.fb_free_pinned, ptr)
    return arr
