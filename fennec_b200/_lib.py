"""ctypes binding of libfennec_b200.so — one prototype per symbol declared in include/fennec_b200.h.

The library is built in-tree by `python -m fennec_b200.build` (nvcc, sm_100a).  There is no CPU
fallback: if the shared object is missing, or it reports no usable GPU, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libfennec_b200.so")

FB_OK, FB_IDENTITY = 0, 1
FB_E_INVALID, FB_E_NOGPU, FB_E_CUDA, FB_E_OOM, FB_E_CANCELLED = -1, -2, -3, -4, -5
FB_OP_SSIM, FB_OP_SSIM_FAST, FB_OP_MSSSIM = 0, 1, 2
FB_FX_GAUSSIAN_BLUR, FB_FX_SHARPEN, FB_FX_ADAPTIVE_SHARPEN = 0, 1, 2

u8p = C.POINTER(C.c_uint8)
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class FbWeights(C.Structure):
    """struct fb_weights (CSR filter taps, resize.go:71-74)."""
    _fields_ = [("n", C.c_int), ("start", ip), ("index", ip), ("weight", dp)]


class FbImageStats(C.Structure):
    """struct fb_image_stats — the fields of fennec.ImageStats (analyze.go:9-22)."""
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("has_alpha", C.c_int), ("is_grayscale", C.c_int),
                ("unique_colors", C.c_int), ("entropy", C.c_double), ("edge_density", C.c_double),
                ("mean_brightness", C.c_double), ("contrast", C.c_double), ("recommended_format", C.c_int),
                ("recommended_quality", C.c_int), ("estimated_compression", C.c_double)]


PROGRESS_FN = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_void_p)


class FbBatchOpts(C.Structure):
    """struct fb_batch_opts (BatchOptions, batch.go:33-44)."""
    _fields_ = [("workers_per_device", C.c_int), ("cancel", ip), ("on_item", PROGRESS_FN), ("user", C.c_void_p)]


class FbPair(C.Structure):
    _fields_ = [("a", C.c_void_p), ("strideA", C.c_int), ("b", C.c_void_p), ("strideB", C.c_int), ("w", C.c_int), ("h", C.c_int)]


class FbResizeItem(C.Structure):
    _fields_ = [("src", C.c_void_p), ("srcStride", C.c_int), ("srcW", C.c_int), ("srcH", C.c_int),
                ("dst", C.c_void_p), ("dstStride", C.c_int), ("dstW", C.c_int), ("dstH", C.c_int)]


class FbEffectItem(C.Structure):
    _fields_ = [("src", C.c_void_p), ("srcStride", C.c_int), ("dst", C.c_void_p), ("dstStride", C.c_int),
                ("w", C.c_int), ("h", C.c_int)]


_IMG = [u8p, C.c_int]
_PAIR = _IMG + _IMG + [C.c_int, C.c_int, dp]
_BATCH_SCORE = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]

# symbol -> (restype, argtypes): must list every function of include/fennec_b200.h
PROTOTYPES = {
    "fb_init": (C.c_int, [ip, C.c_int]),
    "fb_shutdown": (None, []),
    "fb_device_count": (C.c_int, []),
    "fb_set_device": (C.c_int, [C.c_int]),
    "fb_last_error": (C.c_char_p, []),
    "fb_version": (C.c_char_p, []),
    "fb_ssim": (C.c_int, _PAIR),
    "fb_ssim_fast": (C.c_int, _PAIR),
    "fb_msssim": (C.c_int, _PAIR),
    "fb_pixel_ssim": (C.c_int, _PAIR),
    "fb_box_downsample": (C.c_int, _IMG + [C.c_int, C.c_int] + _IMG + [C.c_int, C.c_int]),
    "fb_ssim_fast_dims": (C.c_int, [C.c_int, C.c_int, ip, ip]),
    "fb_gaussian_blur": (C.c_int, _IMG + [C.c_int, C.c_int, dp, C.c_int] + _IMG),
    "fb_blur_kernel": (C.c_int, [C.c_double, dp, C.c_int]),
    "fb_gaussian_blur_sigma": (C.c_int, _IMG + [C.c_int, C.c_int, C.c_double] + _IMG),
    "fb_blur3x3": (C.c_int, _IMG + [C.c_int, C.c_int] + _IMG),
    "fb_sharpen": (C.c_int, _IMG + [C.c_int, C.c_int, C.c_double] + _IMG),
    "fb_adaptive_sharpen": (C.c_int, _IMG + [C.c_int, C.c_int, C.c_double] + _IMG),
    "fb_lanczos_weights_cap": (C.c_int, [C.c_int, C.c_int]),
    "fb_build_lanczos_weights": (C.c_int, [C.c_int, C.c_int, ip, ip, dp]),
    "fb_lanczos_resize": (C.c_int, _IMG + [C.c_int, C.c_int] + _IMG + [C.c_int, C.c_int,
                                                                        C.POINTER(FbWeights), C.POINTER(FbWeights)]),
    "fb_smart_resize_dims": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]),
    "fb_ycbcr_to_nrgba": (C.c_int, [u8p, C.c_int, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int]),
    "fb_gray_to_nrgba": (C.c_int, [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int]),
    "fb_convert_to_nrgba": (C.c_int, [C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint16), C.c_int, u8p, C.c_int]),
    "fb_convert_to_nrgba_batch_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                                C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int]),
    "fb_ssim_ref_create": (C.c_int, _IMG + [C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "fb_ssim_ref_score_ycbcr": (C.c_int, [C.c_void_p, u8p, C.c_int, u8p, u8p, C.c_int, C.c_int, dp]),
    "fb_ssim_ref_score_nrgba": (C.c_int, [C.c_void_p] + _IMG + [dp]),
    "fb_ssim_ref_destroy": (None, [C.c_void_p]),
    "fb_ycbcr_to_nrgba_batch_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int,
                                              C.c_int]),
    "fb_analyze": (C.c_int, _IMG + [C.c_int, C.c_int, C.POINTER(FbImageStats)]),
    "fb_analyze_raw_bytes": (C.c_size_t, []),
    "fb_analyze_batch_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fb_analyze_finish": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(FbImageStats)]),
    "fb_orientation_dims": (C.c_int, [C.c_int, C.c_int, C.c_int, ip, ip]),
    "fb_apply_orientation": (C.c_int, _IMG + [C.c_int, C.c_int, C.c_int] + _IMG),
    "fb_apply_orientation_batch_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_int64, C.c_int, C.c_int]),
    "fb_apply_palette": (C.c_int, _IMG + [C.c_int, C.c_int, u8p, C.c_int, u8p, C.c_int, u8p, C.c_int]),
    "fb_apply_palette_batch_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int]),
    "fb_ssim_batch_dev": (C.c_int, _BATCH_SCORE),
    "fb_ssim_fast_batch_dev": (C.c_int, _BATCH_SCORE),
    "fb_msssim_batch_dev": (C.c_int, _BATCH_SCORE),
    "fb_box_downsample_batch_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]),
    "fb_msssim_level_batch_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "fb_gaussian_blur_batch_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                             C.c_int, C.c_int, dp, C.c_int]),
    "fb_sharpen_batch_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_double]),
    "fb_adaptive_sharpen_batch_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                                C.c_int, C.c_int, C.c_int, C.c_double]),
    "fb_lanczos_resize_batch_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]),
    "fb_workspace_bytes": (C.c_size_t, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "fb_batch_shard": (C.c_int, [C.c_int, C.c_int, C.c_int, ip, ip]),
    "fb_take_launch_count": (C.c_longlong, []),
    "fb_msssim_level2_batch_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "fb_score_batch_host": (C.c_int, [C.c_int, C.POINTER(FbPair), C.c_int, dp, ip, C.POINTER(FbBatchOpts)]),
    "fb_lanczos_resize_batch_host": (C.c_int, [C.POINTER(FbResizeItem), C.c_int, ip, C.POINTER(FbBatchOpts)]),
    "fb_effect_batch_host": (C.c_int, [C.c_int, C.c_double, C.POINTER(FbEffectItem), C.c_int, ip, C.POINTER(FbBatchOpts)]),
    "fb_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "fb_free_pinned": (None, [C.c_void_p]),
    "fb_debug_pool_size": (C.c_int, []),
    "fb_debug_table_count": (C.c_int, []),
}


class FennecError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libfennec_b200 status {status}: {message}")
        self.status = status


_lib = None


def load():
    """Load the shared library (no GPU needed for loading or for the host-only helpers)."""
    global _lib
    if _lib is None:
        path = os.environ.get("FB_LIB_PATH") or SO_PATH   # FB_LIB_PATH: tuning builds of tools/build_variant.sh
        if not os.path.exists(path):
            raise FennecError(FB_E_INVALID, f"{path} not built — run `python -m fennec_b200.build`")
        lib = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the binary drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int) -> int:
    """Raise on negative status; return it otherwise (FB_OK or FB_IDENTITY)."""
    if status < 0:
        raise FennecError(status, load().fb_last_error().decode("utf-8", "replace"))
    return status
