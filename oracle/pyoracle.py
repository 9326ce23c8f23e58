"""ctypes binding of oracle/_build/libfennec_oracle.so (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (fennec_b200) never does.  PARITY UNPINNED — see
fennec_oracle.h.  Images are numpy uint8 arrays of shape (h, w, 4), NRGBA, C-contiguous rows
(row stride = arr.strides[0] bytes, so padded strides are exercised through views).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfennec_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile). Building the checker is not using it."""
    src = os.path.join(_HERE, "fennec_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


class FoImageStats(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("has_alpha", C.c_int), ("is_grayscale", C.c_int),
                ("unique_colors", C.c_int), ("entropy", C.c_double), ("edge_density", C.c_double),
                ("mean_brightness", C.c_double), ("contrast", C.c_double), ("recommended_format", C.c_int),
                ("recommended_quality", C.c_int), ("estimated_compression", C.c_double), ("histogram", C.c_double * 256)]


_lib = None
_u8p = C.POINTER(C.c_uint8)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        img = [_u8p, C.c_int]
        L.fo_set_procs.argtypes = [C.c_int]
        L.fo_get_procs.restype = C.c_int
        L.fo_clampf.argtypes = [C.c_double]
        L.fo_clampf.restype = C.c_uint8
        L.fo_gaussian_kernel.argtypes = [C.c_int, C.c_double, _dp]
        L.fo_to_luminance.argtypes = img + [C.c_int, C.c_int, _dp]
        L.fo_windowed_ssim.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int]
        L.fo_windowed_ssim.restype = C.c_double
        for name in ("fo_pixel_ssim", "fo_ssim", "fo_ssim_fast", "fo_msssim"):
            f = getattr(L, name)
            f.argtypes = img + img + [C.c_int, C.c_int]
            f.restype = C.c_double
        L.fo_ssim_fast_dims.argtypes = [C.c_int, C.c_int, _ip, _ip]
        L.fo_box_downsample.argtypes = img + [C.c_int, C.c_int] + img + [C.c_int, C.c_int]
        L.fo_blur_radius.argtypes = [C.c_double]
        L.fo_blur_kernel.argtypes = [C.c_double, C.c_int, _dp]
        L.fo_gaussian_blur_k.argtypes = img + [C.c_int, C.c_int, _dp, C.c_int] + img
        L.fo_gaussian_blur.argtypes = img + [C.c_int, C.c_int, C.c_double] + img
        L.fo_blur3x3.argtypes = img + [C.c_int, C.c_int] + img
        L.fo_sharpen.argtypes = img + [C.c_int, C.c_int, C.c_double] + img
        L.fo_adaptive_sharpen.argtypes = img + [C.c_int, C.c_int, C.c_double] + img
        L.fo_lanczos_kernel.argtypes = [C.c_double]
        L.fo_lanczos_kernel.restype = C.c_double
        L.fo_lanczos_weights_cap.argtypes = [C.c_int, C.c_int]
        L.fo_lanczos_weights.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp]
        L.fo_resize_h.argtypes = img + [C.c_int, C.c_int] + img + [C.c_int]
        L.fo_resize_v.argtypes = img + [C.c_int, C.c_int] + img + [C.c_int]
        L.fo_lanczos_resize.argtypes = img + [C.c_int, C.c_int] + img + [C.c_int, C.c_int]
        L.fo_smart_resize_dims.argtypes = [C.c_int] * 4 + [_ip, _ip]
        L.fo_analyze.argtypes = img + [C.c_int, C.c_int, C.POINTER(FoImageStats)]
        L.fo_analyze.restype = None
        L.fo_apply_orientation.argtypes = img + [C.c_int, C.c_int, C.c_int] + img
        L.fo_apply_palette.argtypes = img + [C.c_int, C.c_int, _u8p, C.c_int, _u8p, C.c_int] + img
        L.fo_apply_palette.restype = None
        L.fo_ycbcr_to_nrgba.argtypes = [_u8p, C.c_int, _u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
        L.fo_gray_to_nrgba.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
        L.fo_convert_to_nrgba.argtypes = [C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint16), C.c_int, _u8p, C.c_int]
        L.fo_convert_to_nrgba.restype = C.c_int
        _lib = L
    return _lib


def _img(a: np.ndarray):
    assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 4
    assert a.strides[2] == 1 and a.strides[1] == 4, "pixels must be interleaved NRGBA"
    return a.ctypes.data_as(_u8p), int(a.strides[0]) if a.shape[0] > 1 else int(a.shape[1] * 4)


def _new(h: int, w: int) -> np.ndarray:
    return np.zeros((max(h, 0), max(w, 0), 4), dtype=np.uint8)


def set_procs(n: int) -> None:
    lib().fo_set_procs(int(n))


def clampf(x: float) -> int:
    return int(lib().fo_clampf(float(x)))


def gaussian_kernel(size: int = 8, sigma: float = 1.5) -> np.ndarray:
    out = np.zeros(size * size, dtype=np.float64)
    lib().fo_gaussian_kernel(size, sigma, out.ctypes.data_as(_dp))
    return out


def to_luminance(a: np.ndarray) -> np.ndarray:
    h, w = a.shape[:2]
    out = np.zeros((h, w), dtype=np.float64)
    p, s = _img(a)
    lib().fo_to_luminance(p, s, w, h, out.ctypes.data_as(_dp))
    return out


def windowed_ssim(la: np.ndarray, lb: np.ndarray, procs: int = 0) -> float:
    h, w = la.shape
    la = np.ascontiguousarray(la, dtype=np.float64)
    lb = np.ascontiguousarray(lb, dtype=np.float64)
    return float(lib().fo_windowed_ssim(la.ctypes.data_as(_dp), lb.ctypes.data_as(_dp), w, h, procs))


def _pair(fn, a, b):
    assert a.shape == b.shape
    h, w = a.shape[:2]
    pa, sa = _img(a)
    pb, sb = _img(b)
    return float(fn(pa, sa, pb, sb, w, h))


def pixel_ssim(a, b):
    return _pair(lib().fo_pixel_ssim, a, b)


def ssim(a, b):
    return _pair(lib().fo_ssim, a, b)


def ssim_fast(a, b):
    return _pair(lib().fo_ssim_fast, a, b)


def msssim(a, b):
    return _pair(lib().fo_msssim, a, b)


def ssim_fast_dims(w: int, h: int):
    nw, nh = C.c_int(), C.c_int()
    did = lib().fo_ssim_fast_dims(w, h, C.byref(nw), C.byref(nh))
    return bool(did), nw.value, nh.value


def box_downsample(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw = src.shape[:2]
    if sw <= 0 or sh <= 0 or dw <= 0 or dh <= 0:
        return _new(0, 0)
    dst = _new(dh, dw)
    ps, ss = _img(src)
    pd, sd = _img(dst)
    lib().fo_box_downsample(ps, ss, sw, sh, pd, sd, dw, dh)
    return dst


def blur_kernel(sigma: float):
    r = lib().fo_blur_radius(sigma)
    k = np.zeros(2 * r + 1, dtype=np.float64)
    lib().fo_blur_kernel(sigma, r, k.ctypes.data_as(_dp))
    return k, r


def gaussian_blur(src: np.ndarray, sigma: float) -> np.ndarray:
    """Returns `src` itself (same object) for sigma <= 0, like the reference (effects.go:147-149)."""
    if sigma <= 0:
        return src
    h, w = src.shape[:2]
    dst = _new(h, w)
    ps, ss = _img(src)
    pd, sd = _img(dst)
    lib().fo_gaussian_blur(ps, ss, w, h, float(sigma), pd, sd)
    return dst


def blur3x3(src: np.ndarray) -> np.ndarray:
    h, w = src.shape[:2]
    dst = _new(h, w)
    ps, ss = _img(src)
    pd, sd = _img(dst)
    lib().fo_blur3x3(ps, ss, w, h, pd, sd)
    return dst


def _fx(fn, src, strength):
    h, w = src.shape[:2]
    dst = _new(h, w)
    ps, ss = _img(src)
    pd, sd = _img(dst)
    if fn(ps, ss, w, h, float(strength), pd, sd) == 1:
        return src  # identity guards return the same pointer (effects.go:11-22,50-61)
    return dst


def sharpen(src, strength):
    return _fx(lib().fo_sharpen, src, strength)


def adaptive_sharpen(src, strength):
    return _fx(lib().fo_adaptive_sharpen, src, strength)


def lanczos_kernel(x: float) -> float:
    return float(lib().fo_lanczos_kernel(float(x)))


def lanczos_weights(dst_size: int, src_size: int):
    cap = lib().fo_lanczos_weights_cap(dst_size, src_size)
    start = np.zeros(dst_size + 1, dtype=np.int32)
    index = np.zeros(cap + 1, dtype=np.int32)
    weight = np.zeros(cap + 1, dtype=np.float64)
    n = lib().fo_lanczos_weights(dst_size, src_size, start.ctypes.data_as(_ip),
                                 index.ctypes.data_as(_ip), weight.ctypes.data_as(_dp))
    return start, index[:n].copy(), weight[:n].copy()


def resize_h(src: np.ndarray, dw: int) -> np.ndarray:
    sh, sw = src.shape[:2]
    dst = _new(sh, dw)
    ps, ss = _img(src)
    pd, sd = _img(dst)
    lib().fo_resize_h(ps, ss, sw, sh, pd, sd, dw)
    return dst


def resize_v(src: np.ndarray, dh: int) -> np.ndarray:
    sh, sw = src.shape[:2]
    dst = _new(dh, sw)
    ps, ss = _img(src)
    pd, sd = _img(dst)
    lib().fo_resize_v(ps, ss, sw, sh, pd, sd, dh)
    return dst


def lanczos_resize(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw = src.shape[:2]
    if sw <= 0 or sh <= 0 or dw <= 0 or dh <= 0:
        return _new(0, 0)
    dst = _new(dh, dw)
    ps, ss = _img(src)
    pd, sd = _img(dst)
    lib().fo_lanczos_resize(ps, ss, sw, sh, pd, sd, dw, dh)
    return dst


def smart_resize_dims(sw: int, sh: int, max_w: int, max_h: int):
    dw, dh = C.c_int(), C.c_int()
    noop = lib().fo_smart_resize_dims(sw, sh, max_w, max_h, C.byref(dw), C.byref(dh))
    return bool(noop), dw.value, dh.value


def _plane(a: np.ndarray):
    assert a.dtype == np.uint8 and a.ndim == 2 and a.strides[1] == 1
    return a.ctypes.data_as(_u8p), int(a.strides[0])


def ycbcr_to_nrgba(y: np.ndarray, cb: np.ndarray, cr: np.ndarray, ratio: int) -> np.ndarray:
    """convertToNRGBA (convert.go:34-64) of an *image.YCbCr with Rect.Min == (0,0); planes are 2-D uint8 arrays."""
    h, w = y.shape
    assert cb.shape == cr.shape and cb.strides == cr.strides
    dst = _new(h, w)
    py, sy = _plane(y)
    pcb, sc = _plane(cb)
    pcr, _ = _plane(cr)
    pd, sd = _img(dst)
    assert lib().fo_ycbcr_to_nrgba(py, sy, pcb, pcr, sc, w, h, ratio, pd, sd) == 0
    return dst


def gray_to_nrgba(g: np.ndarray) -> np.ndarray:
    h, w = g.shape
    dst = _new(h, w)
    pg, sg = _plane(g)
    pd, sd = _img(dst)
    lib().fo_gray_to_nrgba(pg, sg, w, h, pd, sd)
    return dst


_FMT_BPP = {1: 4, 2: 8, 3: 8, 4: 2, 5: 4, 6: 1}


def convert_to_nrgba(fmt: int, pix: np.ndarray, pal16: np.ndarray = None) -> np.ndarray:
    """convertToNRGBA (convert.go:34-64) of *image.RGBA (1), RGBA64 (2), NRGBA64 (3), Gray16 (4), CMYK (5), Paletted (6).
    pix: (h, w, bytes-per-pixel) uint8 — (h, w) for Paletted; rows may be strided."""
    bpp = _FMT_BPP[fmt]
    if pix.ndim == 2:
        pix = pix[..., None]
    h, w = pix.shape[:2]
    assert pix.dtype == np.uint8 and pix.shape[2] == bpp
    assert (bpp == 1 or pix.strides[2] == 1) and (w <= 1 or pix.strides[1] == bpp), "pixels must be packed within a row"
    dst = _new(h, w)
    pd, sd = _img(dst)
    pp = None
    n = 0
    if pal16 is not None:
        pal16 = np.ascontiguousarray(pal16, dtype=np.uint16)
        pp, n = pal16.ctypes.data_as(C.POINTER(C.c_uint16)), len(pal16)
    stride = int(pix.strides[0]) if h > 1 else w * bpp
    rc = lib().fo_convert_to_nrgba(fmt, pix.ctypes.data_as(_u8p), stride, w, h, pp, n, pd, sd)
    if rc == -2:
        raise IndexError("palette index out of range (Go panics)")
    assert rc == 0
    return dst


def analyze(img: np.ndarray) -> dict:
    """Analyze (analyze.go:26-176) → dict of the ImageStats fields (+ the luminance histogram)."""
    st = FoImageStats()
    if img.size == 0:
        lib().fo_analyze(None, 0, img.shape[1] if img.ndim == 3 else 0, 0, C.byref(st))
    else:
        p, s = _img(img)
        lib().fo_analyze(p, s, img.shape[1], img.shape[0], C.byref(st))
    d = {k: getattr(st, k) for k, _ in FoImageStats._fields_ if k != "histogram"}
    d["histogram"] = np.array(st.histogram[:], dtype=np.float64)
    return d


def apply_orientation(src: np.ndarray, orient: int) -> np.ndarray:
    """ApplyOrientation (exif.go:176-203); identity orientations return `src` itself."""
    h, w = src.shape[:2]
    if orient < 2 or orient > 8:
        return src
    dst = _new(w, h) if orient >= 5 else _new(h, w)
    ps, ss = _img(src)
    pd, sd = _img(dst)
    assert lib().fo_apply_orientation(ps, ss, w, h, orient, pd, sd) == 0
    return dst


def apply_palette(src: np.ndarray, palette: np.ndarray):
    """applyPalette + palettedToNRGBA (targetsize.go:479-545) → (indices (h, w), reconstruction (h, w, 4))."""
    h, w = src.shape[:2]
    pal = np.ascontiguousarray(palette, dtype=np.uint8)
    idx = np.zeros((h, w), np.uint8)
    out = _new(h, w)
    ps, ss = _img(src)
    po, so = _img(out)
    lib().fo_apply_palette(ps, ss, w, h, pal.ctypes.data_as(_u8p), pal.shape[0], idx.ctypes.data_as(_u8p), w, po, so)
    return idx, out
