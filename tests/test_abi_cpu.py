"""CPU suite: the C-ABI library loads, exports every symbol include/fennec_b200.h declares, and its
host-side helpers (weight tables, dimension rules, sharder) agree with the oracle.  No compute call
is made here; without a GPU compute entry points must fail loudly (FB_E_NOGPU), never fall back."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from fennec_b200 import _lib, api, batch
from fennec_b200 import synth as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "fennec_b200.h")).read()
    return sorted(set(re.findall(r"FB_API[^;(]*?\b(fb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = header_symbols()
    assert len(syms) >= 30
    for must in ("fb_ssim", "fb_ssim_fast", "fb_msssim", "fb_box_downsample", "fb_gaussian_blur", "fb_sharpen",
                 "fb_adaptive_sharpen", "fb_lanczos_resize", "fb_batch_shard", "fb_ssim_batch_dev"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert set(header_symbols()) == set(_lib.PROTOTYPES), "ctypes prototypes and header drifted apart"


def test_version_and_error_strings(lib):
    assert b"fennec-b200" in lib.fb_version()
    assert isinstance(lib.fb_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="asserts the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_fallback(lib):
    a = np.zeros((16, 16, 4), dtype=np.uint8)
    with pytest.raises(_lib.FennecError) as e:
        api.SSIM(a, a)
    assert e.value.status == _lib.FB_E_NOGPU
    with pytest.raises(_lib.FennecError):
        api.GaussianBlur(a, 1.0)
    with pytest.raises(_lib.FennecError):
        api.lanczos_resize(a, 8, 8)


def test_identity_guards_need_no_gpu(lib):
    # The guards that return the reference's input pointer are host logic (effects.go:11-22,147-149).
    a = np.zeros((16, 16, 4), dtype=np.uint8)
    assert api.GaussianBlur(a, 0.0) is a and api.GaussianBlur(a, -2.0) is a
    assert api.Sharpen(a, 0.0) is a and api.AdaptiveSharpen(a, -1.0) is a
    tiny = np.zeros((2, 2, 4), dtype=np.uint8)
    assert api.Sharpen(tiny, 0.5) is tiny and api.AdaptiveSharpen(tiny, 0.5) is tiny
    assert api.lanczos_resize(a, 0, 10).shape == (0, 0, 4)      # resize.go:41-43
    assert api.box_downsample(a, 10, -1).shape == (0, 0, 4)     # ssim.go:246-248
    assert api.smart_resize(a, 100, 100) is a                   # resize.go:23-25
    same = api.lanczos_resize(a, 16, 16)                        # resize.go:45-49: a copy, not the pointer
    assert same is not a and np.array_equal(same, a)


def test_lanczos_weight_builder_matches_oracle(lib, oracle, golden):
    for dst, src in ((1920, 7680), (100, 400), (333, 120), (37, 400), (1, 40), (7, 7)):
        st, ix, wt = api.lanczos_weights(dst, src)
        ost, oix, owt = oracle.lanczos_weights(dst, src)
        assert np.array_equal(st, ost) and np.array_equal(ix, oix) and np.array_equal(wt, owt)
    # config 4 geometry: 24 taps interior, identical weights for every interior d
    st, ix, wt = api.lanczos_weights(1920, 7680)
    taps = np.diff(st)
    assert taps[0] == 14 and taps[-1] == 14 and set(taps[3:-3]) == {24}
    assert np.array_equal(wt[st[10]:st[11]], wt[st[1000]:st[1001]])
    assert abs(np.abs(wt[st[10]:st[11]]).sum() - 1.369) < 1e-3


def test_dimension_rules_match_oracle(lib, oracle):
    for w, h in ((4032, 3024), (7680, 4320), (512, 512), (513, 100), (600, 9), (5000, 40), (100, 5000), (8, 8)):
        assert api.ssim_fast_dims(w, h) == oracle.ssim_fast_dims(w, h)
    for args in ((4032, 3024, 1920, 1080), (200, 100, 100, 100), (200, 100, 0, 50), (100, 100, 400, 0),
                 (3, 1000, 2, 2), (7680, 4320, 1920, 1080)):
        dw, dh = C.c_int(), C.c_int()
        noop = lib.fb_smart_resize_dims(*args, C.byref(dw), C.byref(dh))
        assert (bool(noop), dw.value, dh.value) == oracle.smart_resize_dims(*args)


def test_blur_kernel_builder_matches_oracle(lib, oracle):
    for sigma in (0.3, 0.5, 1.0, 2.0, 3.3, 20.0):
        k, r = api.blur_kernel(sigma)
        ok, orr = oracle.blur_kernel(sigma)
        assert r == orr and np.array_equal(k, ok)
    with pytest.raises(ValueError):
        api.blur_kernel(0.0)


def test_batch_shard_partition(lib):
    # static contiguous partition that keeps input order (batch.go:71,108)
    for n, g in ((4096, 8), (1024, 8), (10, 3), (3, 8), (0, 4), (1, 1), (17, 4)):
        seen = []
        for s in range(g):
            b, e = batch.shard_range(n, g, s)
            assert 0 <= b <= e <= n
            seen += list(range(b, e))
        assert seen == list(range(n))
    with pytest.raises(_lib.FennecError):
        batch.shard_range(10, 0, 0)
    with pytest.raises(_lib.FennecError):
        batch.shard_range(10, 4, 4)


def test_workspace_bytes_is_host_logic(lib):
    assert lib.fb_workspace_bytes(b"ssim", 3840, 2160, 0, 0, 32) > 0
    assert lib.fb_workspace_bytes(b"msssim", 7680, 4320, 0, 0, 1) > 2 * 3840 * 2160 * 4
    assert lib.fb_workspace_bytes(b"nonsense", 1, 1, 1, 1, 1) == 0


# ---- SURVEY §8(f1-f4) entry points: argument checking and host logic that run without a GPU ------------------------

def test_new_entry_points_validate_before_touching_the_gpu(lib):
    y = np.zeros((4, 8), np.uint8)
    c = np.zeros((2, 4), np.uint8)
    dst = np.zeros((4, 8, 4), np.uint8)
    u8 = _lib.u8p
    p = lambda a: a.ctypes.data_as(u8)  # noqa: E731
    # unknown subsample ratio / plane stride too small / null planes
    assert lib.fb_ycbcr_to_nrgba(p(y), 8, p(c), p(c), 4, 8, 4, 9, p(dst), 32) == _lib.FB_E_INVALID
    assert b"ratio" in lib.fb_last_error()
    assert lib.fb_ycbcr_to_nrgba(p(y), 7, p(c), p(c), 4, 8, 4, 2, p(dst), 32) == _lib.FB_E_INVALID
    assert lib.fb_ycbcr_to_nrgba(None, 8, p(c), p(c), 4, 8, 4, 2, p(dst), 32) == _lib.FB_E_INVALID
    # palette: 0 or > 256 entries, alpha != 255
    pal = np.full((4, 4), 255, np.uint8)
    idx = np.zeros((4, 8), np.uint8)
    assert lib.fb_apply_palette(p(dst), 32, 8, 4, p(pal), 0, p(idx), 8, None, 0) == _lib.FB_E_INVALID
    assert lib.fb_apply_palette(p(dst), 32, 8, 4, p(pal), 257, p(idx), 8, None, 0) == _lib.FB_E_INVALID
    pal[2, 3] = 254
    assert lib.fb_apply_palette(p(dst), 32, 8, 4, p(pal), 4, p(idx), 8, None, 0) == _lib.FB_E_INVALID
    assert b"alpha" in lib.fb_last_error()
    # convertToNRGBA formats: unknown format, short stride, missing / oversized palette, empty image
    pal16 = np.zeros((4, 4), np.uint16).ctypes.data_as(C.POINTER(C.c_uint16))
    assert lib.fb_convert_to_nrgba(0, p(dst), 32, 8, 4, None, 0, p(dst), 32) == _lib.FB_E_INVALID
    assert b"format" in lib.fb_last_error()
    assert lib.fb_convert_to_nrgba(2, p(dst), 32, 8, 4, None, 0, p(dst), 32) == _lib.FB_E_INVALID      # RGBA64 needs 64 B/row
    assert lib.fb_convert_to_nrgba(6, p(idx), 8, 8, 4, None, 0, p(dst), 32) == _lib.FB_E_INVALID
    assert lib.fb_convert_to_nrgba(6, p(idx), 8, 8, 4, pal16, 257, p(dst), 32) == _lib.FB_E_INVALID
    assert lib.fb_convert_to_nrgba(1, None, 0, 0, 0, None, 0, None, 0) == _lib.FB_OK                   # nothing to do
    # session: null outputs
    assert lib.fb_ssim_ref_create(p(dst), 32, 8, 4, None) == _lib.FB_E_INVALID
    assert lib.fb_ssim_ref_score_nrgba(None, p(dst), 32, None) == _lib.FB_E_INVALID
    lib.fb_ssim_ref_destroy(None)   # a no-op, like free(NULL)


def test_orientation_dims_and_identity(lib):   # exif.go:176-203
    dw, dh = C.c_int(), C.c_int()
    for o in (2, 3, 4):
        assert lib.fb_orientation_dims(o, 30, 20, C.byref(dw), C.byref(dh)) == _lib.FB_OK and (dw.value, dh.value) == (30, 20)
    for o in (5, 6, 7, 8):
        assert lib.fb_orientation_dims(o, 30, 20, C.byref(dw), C.byref(dh)) == _lib.FB_OK and (dw.value, dh.value) == (20, 30)
    for o in (0, 1, 9, -3):   # the reference returns its input
        assert lib.fb_orientation_dims(o, 30, 20, C.byref(dw), C.byref(dh)) == _lib.FB_IDENTITY
        img = np.zeros((20, 30, 4), np.uint8)
        assert api.ApplyOrientation(img, o) is img


def test_analyze_finish_is_host_arithmetic(lib, oracle):
    """fb_analyze_finish turns a raw record into ImageStats on the host: feed it a record assembled from the oracle's
    own histogram and check entropy / mean / recommendations (analyze.go:86, 116-128, 183-232) without any GPU."""
    img = S.gradient_noise_image(333, 217, 4)
    want = oracle.analyze(img)
    rec = np.zeros(int(lib.fb_analyze_raw_bytes()), np.uint8)
    hist = rec[:1024].view(np.uint32)
    hist[:] = want["histogram"].astype(np.uint32)
    lum = 299 * img[..., 0].astype(np.int64) + 587 * img[..., 1].astype(np.int64) + 114 * img[..., 2].astype(np.int64)
    rec[1024:1032].view(np.uint64)[0] = int(lum.sum())
    sy, sx = int(np.ceil(217 / 100)), int(np.ceil(333 / 100))
    n = len(range(0, 217, sy)) * len(range(0, 333, sx))
    rec[1032:1040].view(np.float64)[0] = want["contrast"] ** 2 * n          # varSum
    tail = rec[1040:1056].view(np.uint32)                                     # hasAlpha, hasColour, uniqueSampled, edges
    ex, ey = max(1, 333 // 200), max(1, 217 // 200)
    total = len(range(1, 216, ey)) * len(range(1, 332, ex))
    tail[:] = [want["has_alpha"], 1 - want["is_grayscale"], want["unique_colors"], round(want["edge_density"] * total)]
    st = _lib.FbImageStats()
    assert lib.fb_analyze_finish(rec.ctypes.data, 333, 217, C.byref(st)) == _lib.FB_OK
    assert (st.width, st.height, st.has_alpha, st.is_grayscale, st.unique_colors) == (333, 217, want["has_alpha"], want["is_grayscale"], want["unique_colors"])
    assert abs(st.entropy - want["entropy"]) <= 1e-12 and abs(st.mean_brightness - want["mean_brightness"]) <= 1e-9
    assert abs(st.contrast - want["contrast"]) <= 1e-9 and abs(st.edge_density - want["edge_density"]) <= 1e-12
    assert (st.recommended_format, st.recommended_quality) == (want["recommended_format"], want["recommended_quality"])
    assert abs(st.estimated_compression - want["estimated_compression"]) <= 1e-12


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: include/fennec_b200.h must compile as C99 (cgo compiles it as C), not only as C++."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "fennec_b200.h"\nint main(void) { fb_image_stats s; (void)s; return FB_OK; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-c", str(src), "-o", str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
