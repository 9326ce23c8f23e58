// fennec.hpp — C++ host-side mirror of fennec's Go API for the hot path, over the C ABI (fennec_b200.h).
//
// The reference is compiled Go; no Go toolchain exists in the build image, so the host layer above the
// C ABI is provided in C++ (this header) and in Python (fennec_b200/api.py).  Names, argument meaning and
// guard behaviour follow the Go functions (file:line cited per function): where Go returns its input
// pointer unchanged, the SAME shared_ptr is returned; where it returns an empty image, an empty NRGBA.
// Every function calls libfennec_b200.so; there is no CPU compute here.  Errors (no GPU, OOM, bad
// arguments) throw fennec::Error — the Go shim falls back to the pure-Go body instead.
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "fennec_b200.h"

namespace fennec {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &m) : std::runtime_error("libfennec_b200 status " + std::to_string(s) + ": " + m), status(s) {}
};

inline int check(int status) {
    if (status < 0) throw Error(status, fb_last_error());
    return status;
}

// image.NRGBA: Pix (R,G,B,A interleaved, non-premultiplied), Stride bytes per row.
struct NRGBA {
    std::vector<uint8_t> Pix;
    int Stride = 0, W = 0, H = 0;
    NRGBA() = default;
    NRGBA(int w, int h) : Pix((size_t)(w > 0 ? w : 0) * (h > 0 ? h : 0) * 4, 0), Stride((w > 0 ? w : 0) * 4), W(w > 0 ? w : 0), H(h > 0 ? h : 0) {}
    uint8_t *at(int x, int y) { return Pix.data() + (size_t)y * Stride + (size_t)x * 4; }
    const uint8_t *data() const { return Pix.empty() ? nullptr : Pix.data(); }
    uint8_t *data() { return Pix.empty() ? nullptr : Pix.data(); }
};
using Image = std::shared_ptr<NRGBA>;
inline Image NewNRGBA(int w, int h) { return std::make_shared<NRGBA>(w, h); }  // zero-filled like image.NewNRGBA

// lanczosResize — resize.go:37-53
inline Image lanczosResize(const Image &img, int dstW, int dstH) {
    if (img->W <= 0 || img->H <= 0 || dstW <= 0 || dstH <= 0) return NewNRGBA(0, 0);
    Image dst = NewNRGBA(dstW, dstH);
    check(fb_lanczos_resize(img->data(), img->Stride, img->W, img->H, dst->data(), dst->Stride, dstW, dstH, nullptr, nullptr));
    return dst;
}

// smartResize — resize.go:12-32 (returns img itself when it already fits)
inline Image smartResize(const Image &img, int maxW, int maxH) {
    int dw = 0, dh = 0;
    if (fb_smart_resize_dims(img->W, img->H, maxW, maxH, &dw, &dh) == 1) return img;
    return lanczosResize(img, dw, dh);
}

namespace detail {
typedef int (*score_fn)(const uint8_t *, int, const uint8_t *, int, int, int, double *);
inline double score(score_fn fn, const Image &a, Image b, bool resize) {
    if (resize && (b->W != a->W || b->H != a->H)) b = lanczosResize(b, a->W, a->H);  // ssim.go:31-33 / 320-322
    if (b->W != a->W || b->H != a->H) throw Error(FB_E_INVALID, "images must have equal dimensions");
    double out = 0.0;
    check(fn(a->data(), a->Stride, b->data(), b->Stride, a->W, a->H, &out));
    return out;
}
}  // namespace detail

inline double SSIM(const Image &a, const Image &b) { return detail::score(fb_ssim, a, b, true); }          // ssim.go:24-43
inline double SSIMFast(const Image &a, const Image &b) { return detail::score(fb_ssim_fast, a, b, false); }  // ssim.go:48-70
inline double MSSSIM(const Image &a, const Image &b) { return detail::score(fb_msssim, a, b, true); }       // ssim.go:313-365

// boxDownsample — ssim.go:244-309
inline Image boxDownsample(const Image &img, int dstW, int dstH) {
    if (img->W <= 0 || img->H <= 0 || dstW <= 0 || dstH <= 0) return NewNRGBA(0, 0);
    Image dst = NewNRGBA(dstW, dstH);
    check(fb_box_downsample(img->data(), img->Stride, img->W, img->H, dst->data(), dst->Stride, dstW, dstH));
    return dst;
}

// GaussianBlur — effects.go:146-220 (sigma <= 0 → the same pointer)
inline Image GaussianBlur(const Image &img, double sigma) {
    if (sigma <= 0) return img;
    Image dst = NewNRGBA(img->W, img->H);
    check(fb_gaussian_blur_sigma(img->data(), img->Stride, img->W, img->H, sigma, dst->data(), dst->Stride));
    return dst;
}

namespace detail {
typedef int (*fx_fn)(const uint8_t *, int, int, int, double, uint8_t *, int);
inline Image fx(fx_fn fn, const Image &img, double strength) {
    if (strength <= 0 || img->W < 3 || img->H < 3) return img;  // effects.go:11-22 / 50-61
    Image dst = NewNRGBA(img->W, img->H);
    if (check(fn(img->data(), img->Stride, img->W, img->H, strength, dst->data(), dst->Stride)) == FB_IDENTITY) return img;
    return dst;
}
}  // namespace detail

inline Image Sharpen(const Image &img, double strength) { return detail::fx(fb_sharpen, img, strength); }                  // effects.go:10-45
inline Image AdaptiveSharpen(const Image &img, double strength) { return detail::fx(fb_adaptive_sharpen, img, strength); }  // effects.go:49-90

// ---- the callers / data formats either side of the path --------------------------------

// image.YCbCr with Rect.Min == (0,0): what jpeg.Decode returns (image/ycbcr.go).
struct YCbCr {
    std::vector<uint8_t> Y, Cb, Cr;
    int YStride = 0, CStride = 0, W = 0, H = 0;
    int SubsampleRatio = 0;   // image.YCbCrSubsampleRatio: 0 = 444, 1 = 422, 2 = 420, 3 = 440, 4 = 411, 5 = 410
};

// convertToNRGBA — convert.go:34-64 — for a decoded *image.YCbCr
inline Image convertToNRGBA(const YCbCr &img) {
    Image dst = NewNRGBA(img.W, img.H);
    if (img.W <= 0 || img.H <= 0) return dst;
    check(fb_ycbcr_to_nrgba(img.Y.data(), img.YStride, img.Cb.data(), img.Cr.data(), img.CStride, img.W, img.H,
                            img.SubsampleRatio, dst->data(), dst->Stride));
    return dst;
}

// convertToNRGBA — convert.go:34-64 — for the other decoder outputs: *image.RGBA / RGBA64 / NRGBA64 / Gray16 / CMYK /
// Paletted, handed over as the Go image's Pix + Stride (Rect.Min == (0,0)).  Format = FB_FMT_*; Palette16 holds
// Palette[i].RGBA() (4 x uint16 per entry) for FB_FMT_PALETTED.
struct PixImage {
    int Format = FB_FMT_RGBA;
    std::vector<uint8_t> Pix;
    int Stride = 0, W = 0, H = 0;
    std::vector<uint16_t> Palette16;
};
inline Image convertToNRGBA(const PixImage &img) {
    Image dst = NewNRGBA(img.W, img.H);
    if (img.W <= 0 || img.H <= 0) return dst;
    check(fb_convert_to_nrgba(img.Format, img.Pix.data(), img.Stride, img.W, img.H,
                              img.Palette16.empty() ? nullptr : img.Palette16.data(), (int)(img.Palette16.size() / 4),
                              dst->data(), dst->Stride));
    return dst;
}

// The `src` side of compress.go:45-74's binary search kept on the device: SSIMFast(src, candidate) per iteration.
class SSIMSession {
    fb_ssim_ref *h_ = nullptr;
public:
    explicit SSIMSession(const Image &src) { check(fb_ssim_ref_create(src->data(), src->Stride, src->W, src->H, &h_)); }
    ~SSIMSession() { fb_ssim_ref_destroy(h_); }
    SSIMSession(const SSIMSession &) = delete;
    SSIMSession &operator=(const SSIMSession &) = delete;
    double score(const YCbCr &d) const {
        double out = 0.0;
        check(fb_ssim_ref_score_ycbcr(h_, d.Y.data(), d.YStride, d.Cb.data(), d.Cr.data(), d.CStride, d.SubsampleRatio, &out));
        return out;
    }
    double score(const Image &d) const {
        double out = 0.0;
        check(fb_ssim_ref_score_nrgba(h_, d->data(), d->Stride, &out));
        return out;
    }
};

// Analyze — analyze.go:26-113 (ImageStats as fb_image_stats; Format / Quality carry Go's numeric values)
inline fb_image_stats Analyze(const Image &img) {
    fb_image_stats st;
    check(fb_analyze(img->data(), img->Stride, img->W, img->H, &st));
    return st;
}

// ApplyOrientation — exif.go:176-203 (orientations 1, 0 and unknown values return the same pointer)
inline Image ApplyOrientation(const Image &img, int orient) {
    int dw = 0, dh = 0;
    if (check(fb_orientation_dims(orient, img->W, img->H, &dw, &dh)) == FB_IDENTITY) return img;
    Image dst = NewNRGBA(dw, dh);
    check(fb_apply_orientation(img->data(), img->Stride, img->W, img->H, orient, dst->data(), dst->Stride));
    return dst;
}

// applyPalette + palettedToNRGBA — targetsize.go:479-545. palette: NRGBA entries with A == 255 (medianCut's output).
struct Paletted {
    std::vector<uint8_t> Pix;   // one index per pixel, Stride == W
    int W = 0, H = 0;
};
inline Paletted applyPalette(const Image &src, const std::vector<uint8_t> &palette, Image *recon = nullptr) {
    Paletted out;
    out.W = src->W; out.H = src->H;
    out.Pix.assign((size_t)src->W * src->H, 0);
    if (recon) *recon = NewNRGBA(src->W, src->H);
    check(fb_apply_palette(src->data(), src->Stride, src->W, src->H, palette.data(), (int)(palette.size() / 4), out.Pix.data(), src->W,
                           recon ? (*recon)->data() : nullptr, recon ? (*recon)->Stride : 0));
    return out;
}

// The sharder part of CompressBatch (batch.go:58-128): items [begin, end) of shard `shard`.
struct ShardRange { int begin, end; };
inline ShardRange BatchShard(int nItems, int nShards, int shard) {
    ShardRange r{0, 0};
    check(fb_batch_shard(nItems, nShards, shard, &r.begin, &r.end));
    return r;
}

}  // namespace fennec
