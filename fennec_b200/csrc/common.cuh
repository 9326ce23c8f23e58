// common.cuh — runtime plumbing shared by every translation unit of libfennec_b200.so:
// per-thread / per-device streams and scratch arenas, error text, launch counting, and the
// device-side rounding helpers that make uint8 outputs bit-exact with the reference's
// clampF (convert.go:149-158).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/fennec_b200.h"

namespace fb {

// ---- error handling ------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define FB_CUDA(call)                                                         \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return fb::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define FB_TRY(call)                   \
    do {                               \
        int _s = (call);               \
        if (_s < 0) return _s;         \
    } while (0)

// ---- launch accounting (bench.py's gpu_launches) --------------------------------------------
extern thread_local long long t_launches;
#define FB_LAUNCHED(n) (fb::t_launches += (n))

// ---- per-thread, per-device context ----------------------------------------------------------
struct Arena {
    char *base = nullptr;
    size_t cap = 0;
    size_t off = 0;
    void reset() { off = 0; }
    // 256-byte aligned carve; nullptr when the reservation was too small (programming error).
    void *take(size_t bytes) {
        size_t a = (off + 255) & ~size_t(255);
        if (a + bytes > cap) return nullptr;
        off = a + bytes;
        return base + a;
    }
};

struct DevCtx {
    int dev = -1;                    // PHYSICAL CUDA device of the stream and arenas
    cudaStream_t stream = nullptr;  // owned stream used by the host entry points
    Arena ws;                        // device scratch (grow-only)
    Arena pin;                       // pinned host scratch (scores, small tables)
    cudaEvent_t ev = nullptr;        // last async H2D out of the pinned arena
    bool pinBusy = false;
    // Ordering of the device arena between streams: the last stream that was handed arena memory and an event
    // recorded behind its work (api.cu: ApiScope / reserve).
    cudaEvent_t useEv = nullptr;
    cudaStream_t lastStream = nullptr;
    bool usePending = false;
    // Staging for pageable caller buffers (api.cu: upload / download): two pinned chunks and their events.
    char *stage = nullptr;
    cudaEvent_t stageEv[2] = {nullptr, nullptr};
    bool stageBusy[2] = {false, false};
    unsigned stageNext = 0;
    // Side stream of the pipelined Lanczos batch (api.cu: resize_on_device): the vertical pass of one sub-batch runs on it
    // while the horizontal pass of the next runs on the caller's stream; created on first use.
    cudaStream_t side = nullptr;
    cudaEvent_t pipeEv[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t joinEv = nullptr;
};

int ensure_init();
int device_count();
int current_device();  // thread's device for host entry points
// Context of this thread for `dev`; creates stream/events on first use. nullptr on failure.
DevCtx *ctx(int dev);
// Make sure the arenas can hold the given bytes (may synchronise + reallocate).
int reserve(DevCtx *c, cudaStream_t s, size_t dev_bytes, size_t pinned_bytes);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// Row pitch the library uses for its own device copies: 16-byte aligned rows enable 128-bit loads.
static inline int dev_pitch(int w) { return (int)align_up((size_t)w * 4, 16); }

// ---- kernels' host-side launchers (one per .cu) ---------------------------------------------
// All take device pointers and enqueue on `s`; scratch comes from c->ws (already reserved).

// ssim.cu
size_t ssim_scratch_bytes(int w, int h, int n);
int launch_ssim(DevCtx *c, cudaStream_t s, const uint8_t *a, const uint8_t *b, long long imgStrideA,
                long long imgStrideB, int rowStrideA, int rowStrideB, int w, int h, int n,
                double *scores, long long scoreStride, void *scratch);
int launch_pixel_ssim(cudaStream_t s, const uint8_t *a, const uint8_t *b, long long imgStrideA,
                      long long imgStrideB, int rowStrideA, int rowStrideB, int w, int h, int n,
                      double *scores, long long scoreStride);
// MSSSIM tail: per-level scores are combined on the device:
int launch_msssim_combine(cudaStream_t s, const double *levelScores, int nLevels, int n,
                          const double *weights_dev, double *out);

// box.cu
int launch_box(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int srcW,
               int srcH, uint8_t *dst, long long dstImgStride, int dstRowStride, int dstW, int dstH, int n,
               const int *reserved);
int launch_box_pair(cudaStream_t s, const uint8_t *srcA, long long srcImgStrideA, int srcRowStrideA, const uint8_t *srcB,
                    long long srcImgStrideB, int srcRowStrideB, int srcW, int srcH, uint8_t *dstA, uint8_t *dstB,
                    long long dstImgStride, int dstRowStride, int dstW, int dstH, int n);
void box_edges_host(int src, int dst, int *lo, int *hi);
// MS-SSIM level step: thumbnail + half-resolution image of both batches from one read; returns 1 when not applicable.
int launch_box_fused(cudaStream_t s, const uint8_t *srcA, long long srcImgStrideA, int srcRowStrideA,
                     const uint8_t *srcB, long long srcImgStrideB, int srcRowStrideB, int srcW, int srcH,
                     uint8_t *thumbA, uint8_t *thumbB, long long thumbImgStride, int thumbRowStride, int tw, int th,
                     uint8_t *halfA, uint8_t *halfB, long long halfImgStride, int halfRowStride, int n);

// Two levels from one read (box_fused2_kernel): thumbnails of levels l and l+1 and the level-(l+2) image; 1 = not applicable.
int launch_box_fused2(cudaStream_t s, const uint8_t *srcA, long long srcImgStrideA, int srcRowStrideA, const uint8_t *srcB,
                      long long srcImgStrideB, int srcRowStrideB, int srcW, int srcH, uint8_t *thumb0A, uint8_t *thumb0B,
                      long long thumb0ImgStride, int thumb0RowStride, int tw0, int th0, uint8_t *thumb1A, uint8_t *thumb1B,
                      long long thumb1ImgStride, int thumb1RowStride, int tw1, int th1, uint8_t *l2A, uint8_t *l2B,
                      long long l2ImgStride, int l2RowStride, int n);

// effects.cu
int launch_gaussian_blur(cudaStream_t s, const uint8_t *src, uint8_t *dst, long long imgStride,
                         int rowStride, int w, int h, int n, const double *kernel_dev,
                         const float *kernel32_dev, const float *kernel32_host, int radius, double wabs, uint8_t *tmp,
                         long long tmpImgStride, int tmpRowStride);
int launch_blur3x3(cudaStream_t s, const uint8_t *src, uint8_t *dst, long long imgStride, int rowStride,
                   int w, int h, int n, long long dstImgStride, int dstRowStride);
int launch_sharpen(cudaStream_t s, const uint8_t *src, uint8_t *dst, long long imgStride, int rowStride,
                   int w, int h, int n, long long dstImgStride, int dstRowStride, double amount, int adaptive);

// ycbcr.cu — SURVEY §8(f1): convertToNRGBA (convert.go:34-64) for decoded *image.YCbCr / *image.Gray
bool ycbcr_ratio_shifts(int ratio, int *xShift, int *yShift);
int launch_ycbcr_to_nrgba(cudaStream_t s, const uint8_t *y, long long yImgStride, int yStride, const uint8_t *cb,
                          const uint8_t *cr, long long cImgStride, int cStride, int w, int h, int ratio, uint8_t *dst,
                          long long dstImgStride, int dstRowStride, int n);
int launch_gray_to_nrgba(cudaStream_t s, const uint8_t *g, long long gImgStride, int gStride, int w, int h, uint8_t *dst,
                         long long dstImgStride, int dstRowStride, int n);

// analyze.cu — SURVEY §8(f2): the scans behind Analyze (analyze.go:26-176)
constexpr int kAnalyzeTableSlots = 1 << 18;   // 64-bit slots of the sampled-colour table (<= 100 001 samples)
struct AnalyzeRaw {                            // what the device produces per image; finished on the host
    unsigned int hist[256];                    // luminance histogram, bins int(lum + 0.5)
    unsigned long long sumL;                   // sum of 299R + 587G + 114B over all pixels (exact)
    double varSum;                             // sum((lum - mean)^2) over the contrast grid
    unsigned int hasAlpha, hasColour, uniqueSampled, edges;
};
struct AnalyzeSteps {
    int contrastX, contrastY, contrastNx, contrastNy;
    int edgeX, edgeY, edgeNx, edgeNy;
    long long sampleStep;
    int nSamples;
};
void analyze_steps(int w, int h, AnalyzeSteps *s);
size_t analyze_scratch_bytes(int w, int h, int n);
int launch_analyze(cudaStream_t s, const uint8_t *imgs, long long imgStride, int rowStride, int w, int h, int n,
                   AnalyzeRaw *raw, void *scratch);

// orient.cu — SURVEY §8(f4): ApplyOrientation (exif.go:176-203, convert.go:186-256)
bool orient_dims(int orient, int w, int h, int *dw, int *dh);
int launch_orient(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int w, int h, int orient,
                  uint8_t *dst, long long dstImgStride, int dstRowStride, int n);

// palette.cu — SURVEY §8(f3): applyPalette + palettedToNRGBA (targetsize.go:479-545)
int launch_apply_palette(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int w, int h,
                         const uint8_t *palettes_dev, int ncolors, uint8_t *idx, long long idxImgStride, int idxRowStride,
                         uint8_t *out, long long outImgStride, int outRowStride, int n, void *scratch);
size_t palette_scratch_bytes(int w, int h, int n);
int pixfmt_bytes_per_pixel(int fmt);
int launch_pixfmt_to_nrgba(cudaStream_t s, int fmt, const uint8_t *src, long long srcImgStride, int srcRowStride, int w, int h,
                           const uint16_t *pal16_dev, int ncolors, uint8_t *dst, long long dstImgStride, int dstRowStride, int n,
                           unsigned int *badIndex_dev);

// resize.cu
// When srcSize == ratio * dstSize every interior destination shares one weight vector: see resize.cu.
struct IntRatioInfo {
    int ratio = 0, taps = 0, off = 0, dLo = 0, dHi = 0;
    float wsum = 0.f;
    float w[28];
    // fully opaque windows (resize.cu int_ratio_window): normalised weights, their FP32 error bound, the alpha byte
    float wn[28];
    float Eo = 0.f;
    unsigned int opaqueA = 0;   // alpha byte << 24; 0 disables the shortcut
    int wdExact = 0;            // the interior destinations' binary64 weight rows are bit-identical (wd[] below)
    double wd[24];
};
int launch_resize_h(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int srcW,
                    int srcH, uint8_t *dst, long long dstImgStride, int dstRowStride, int dstW, int n,
                    const int *start_dev, const int *index_dev, const double *weight_dev,
                    const float *weight32_dev, int maxTaps, double wabs, const int *first_dev,
                    const float *wpadT_dev, int groups, const IntRatioInfo *ir);
int launch_resize_v(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int srcW,
                    int srcH, uint8_t *dst, long long dstImgStride, int dstRowStride, int dstH, int n,
                    const int *start_dev, const int *index_dev, const double *weight_dev,
                    const float *weight32_dev, int maxTaps, double wabs, const IntRatioInfo *ir);

#ifdef __CUDACC__
// ---- device helpers ---------------------------------------------------------------------------

// clampF (convert.go:149-158): uint8(clamp(int64(math.Round(x)), 0, 255)), Round = half away from 0.
// Exact: t = trunc(x) and x - t are both exact in binary64 for |x| < 2^52.
__device__ __forceinline__ uint32_t clampf_dev(double x) {
    if (!(x > -0.5)) return 0u;      // rounds to <= 0 (also NaN → 0, unreachable on this path)
    if (x >= 254.5) return 255u;
    double t = trunc(x);
    int v = (int)t;
    if (x - t >= 0.5) v += 1;
    return (uint32_t)v;
}

__device__ __forceinline__ uint32_t ld_nc_u32(const uint8_t *p) {
    return __ldg(reinterpret_cast<const uint32_t *>(p));
}
__device__ __forceinline__ uint4 ld_nc_u128(const uint8_t *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_nc_u64(const uint8_t *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
#endif

}  // namespace fb
