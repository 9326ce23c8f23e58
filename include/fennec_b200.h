/*
 * fennec_b200.h — C ABI of libfennec_b200.so: the B200-native replacement for fennec's dense
 * per-pixel hot path (SSIM family, box downsample, blur / sharpen, Lanczos-3 resize, batch sharder).
 *
 * The reference (shamspias/fennec, pure Go) has no FFI; the drop-in boundary is therefore the set
 * of Go function bodies named beside each entry point below (file:line under the reference tree).
 * A Go maintainer keeps every exported signature and guard and replaces the loop bodies with cgo
 * calls to these functions — see INTEGRATION.md for the exact stub.
 *
 * Conventions
 *  - Images are NRGBA: 8-bit interleaved R,G,B,A, non-premultiplied, `stride` bytes per row,
 *    pixel (x,y) at pix[y*stride + x*4] — Go's image.NRGBA{Pix,Stride} (image.NRGBA.PixOffset).
 *  - "Host" entry points take host pointers; the library copies in, runs the CUDA kernels and copies
 *    out before returning (synchronous, like the Go functions they replace).  The caller owns every
 *    buffer; no host pointer is retained after return (cgo pointer-passing rule).
 *  - "_dev" entry points take device pointers valid on `device` and a cudaStream_t passed as void*;
 *    they only enqueue work (no synchronisation) so callers can time them with events on that stream.
 *  - Return value: FB_OK (0); FB_IDENTITY (1) when the reference would return its INPUT pointer
 *    unchanged or an empty image (dst is not written); negative = error, fb_last_error() has the text
 *    for the calling thread.  There is NO CPU fallback inside the library: without a usable GPU every
 *    compute entry point returns FB_E_NOGPU.  (The Go shim falls back to the pure-Go body.)
 *  - Thread safety: every entry point may be called concurrently (CompressBatch workers,
 *    batch.go:84-124); each calling thread gets its own stream and workspace per device.
 *  - Size limits: the entries are written for images up to 65535 pixels per side (JPEG's own limit); some map
 *    image rows to grid rows and return FB_E_CUDA beyond that instead of computing — never a wrong result.
 *  - Floating point: scores are binary64; SSIM/MS-SSIM agree with the reference to <= 1e-5 absolute
 *    (measured <= 2e-6); every uint8 output is bit-exact with the reference's FP64 arithmetic.
 */
#ifndef FENNEC_B200_H
#define FENNEC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FB_API __attribute__((visibility("default")))
#else
#define FB_API
#endif

#define FB_OK 0
#define FB_IDENTITY 1
#define FB_E_INVALID (-1)
#define FB_E_NOGPU (-2)
#define FB_E_CUDA (-3)
#define FB_E_OOM (-4)
#define FB_E_CANCELLED (-5) /* batch item not started: the caller's cancel flag was raised (batch.go:90-98) */

/* ---- lifecycle --------------------------------------------------------------------------- */

/* Select the devices this process uses (n == 0: all visible). Idempotent. Returns the device count
 * or a negative error. Compute entry points call it lazily with n == 0. */
FB_API int fb_init(const int *devices, int n);
FB_API void fb_shutdown(void);
FB_API int fb_device_count(void);
/* Device used by host entry points called from THIS thread (default 0). */
FB_API int fb_set_device(int device);
FB_API const char *fb_last_error(void);
FB_API const char *fb_version(void);

/* ---- SSIM family (ssim.go) --------------------------------------------------------------- */

/* fennec.SSIM for equal-sized NRGBA inputs — ssim.go:24-43 (the Lanczos pre-resize of `b` at
 * ssim.go:31-33 is fb_lanczos_resize, called by the host side first). w<8||h<8 → pixelSSIM. */
FB_API int fb_ssim(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h, double *out);
/* fennec.SSIMFast — ssim.go:48-70: box-downsample to <=512, then the windowed kernel. */
FB_API int fb_ssim_fast(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h, double *out);
/* fennec.MSSSIM for equal-sized inputs — ssim.go:313-365. */
FB_API int fb_msssim(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h, double *out);
/* pixelSSIM — ssim.go:169-204 (global statistics; any size). */
FB_API int fb_pixel_ssim(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h, double *out);
/* boxDownsample — ssim.go:244-309. FB_IDENTITY when any dim <= 0 (reference returns an empty image). */
FB_API int fb_box_downsample(const uint8_t *src, int srcStride, int srcW, int srcH,
                      uint8_t *dst, int dstStride, int dstW, int dstH);
/* The dims SSIMFast downsamples to — ssim.go:52-56. Returns 1 if a downsample happens, else 0. */
FB_API int fb_ssim_fast_dims(int w, int h, int *newW, int *newH);

/* ---- effects (effects.go) ---------------------------------------------------------------- */

/* GaussianBlur — effects.go:146-220 with the 1-D kernel supplied by the caller (2*radius+1 weights,
 * built by the host with ITS libm: Go's math.Exp on the Go side — SURVEY.md H5). src != dst. */
FB_API int fb_gaussian_blur(const uint8_t *src, int srcStride, int w, int h,
                     const double *kernel, int radius, uint8_t *dst, int dstStride);
/* Convenience builder for that kernel — effects.go:153-165 (glibc exp). Returns the radius, writes
 * 2*radius+1 weights if cap is large enough (else returns -(needed entries)). */
FB_API int fb_blur_kernel(double sigma, double *kernel, int cap);
/* GaussianBlur with the kernel built inside (sigma <= 0 → FB_IDENTITY, effects.go:147-149). */
FB_API int fb_gaussian_blur_sigma(const uint8_t *src, int srcStride, int w, int h, double sigma,
                           uint8_t *dst, int dstStride);
/* gaussianBlur3x3 — effects.go:116-141. */
FB_API int fb_blur3x3(const uint8_t *src, int srcStride, int w, int h, uint8_t *dst, int dstStride);
/* Sharpen — effects.go:10-45; AdaptiveSharpen — effects.go:49-112. FB_IDENTITY for the guards
 * (strength <= 0, w < 3, h < 3: the reference returns the same pointer). */
FB_API int fb_sharpen(const uint8_t *src, int srcStride, int w, int h, double strength,
               uint8_t *dst, int dstStride);
FB_API int fb_adaptive_sharpen(const uint8_t *src, int srcStride, int w, int h, double strength,
                        uint8_t *dst, int dstStride);

/* ---- Lanczos-3 resize (resize.go) -------------------------------------------------------- */

/* Per-destination-index filter taps in CSR form — the [][]weightEntry of resize.go:71-74,164-197.
 * Taps of destination d are index[start[d] .. start[d+1]) with weight[...]; start has n+1 entries. */
typedef struct fb_weights {
    int n;
    const int *start;
    const int *index;
    const double *weight;
} fb_weights;

/* precomputeWeights — resize.go:164-197 (glibc sin). cap() bounds the entry count. */
FB_API int fb_lanczos_weights_cap(int dstSize, int srcSize);
FB_API int fb_build_lanczos_weights(int dstSize, int srcSize, int *start, int *index, double *weight);
/* lanczosResize — resize.go:37-53 (resizeH :77-118 then resizeV :121-161, uint8 intermediate).
 * wx / wy may be NULL: the library then builds the tables itself. FB_IDENTITY for any dim <= 0
 * (reference returns an empty image). Same dims → plain copy (resize.go:45-49). */
FB_API int fb_lanczos_resize(const uint8_t *src, int srcStride, int srcW, int srcH,
                      uint8_t *dst, int dstStride, int dstW, int dstH,
                      const fb_weights *wx, const fb_weights *wy);
/* smartResize's dimension logic — resize.go:12-32. Returns 1 when it is a no-op (same pointer). */
FB_API int fb_smart_resize_dims(int srcW, int srcH, int maxW, int maxH, int *dstW, int *dstH);

/* ---- the step before SSIMFast in the quality search (compress.go:53-62) --------- */

/* convertToNRGBA — convert.go:34-64 — for the concrete types jpeg.Decode returns, Rect.Min == (0,0).
 * `ratio` is Go's image.YCbCrSubsampleRatio constant: 0 = 4:4:4, 1 = 4:2:2, 2 = 4:2:0, 3 = 4:4:0,
 * 4 = 4:1:1, 5 = 4:1:0 (chroma planes hold ceil(w/2^xs) x ceil(h/2^ys) samples, stride cStride).
 * Bytes are those of img.At(x,y).RGBA() >> 8 (Go's color.YCbCr.RGBA integer transform), alpha 255. */
FB_API int fb_ycbcr_to_nrgba(const uint8_t *y, int yStride, const uint8_t *cb, const uint8_t *cr, int cStride,
                      int w, int h, int ratio, uint8_t *dst, int dstStride);
/* The same for *image.Gray (grayscale JPEGs). */
FB_API int fb_gray_to_nrgba(const uint8_t *g, int gStride, int w, int h, uint8_t *dst, int dstStride);

/* convertToNRGBA — convert.go:34-64 — for the other concrete types image/png and image/jpeg decode into.  `pix` /
 * `stride` are the Go image's Pix / Stride (Rect.Min == (0,0)); the result is byte-for-byte what the reference's
 * img.At(x,y).RGBA() loop writes, including the un-premultiply of translucent pixels (convert.go:54-59) and Go's
 * truncating uint8() for colours above their alpha.
 *   FB_FMT_RGBA     *image.RGBA     4 B/px R,G,B,A premultiplied      FB_FMT_RGBA64   *image.RGBA64   8 B/px big-endian
 *   FB_FMT_NRGBA64  *image.NRGBA64  8 B/px big-endian, straight alpha FB_FMT_GRAY16   *image.Gray16   2 B/px big-endian
 *   FB_FMT_CMYK     *image.CMYK     4 B/px C,M,Y,K                    FB_FMT_PALETTED *image.Paletted 1 B/px index
 * For FB_FMT_PALETTED `palette16` holds ncolors (1..256) entries of 4 uint16 — Palette[i].RGBA() as Go returns it
 * (the Go shim evaluates the color.Color interface once per entry); other formats ignore it (NULL).  Go panics on
 * an index >= len(Palette): the host entry writes such pixels as 0 and returns FB_E_INVALID. */
#define FB_FMT_RGBA 1
#define FB_FMT_RGBA64 2
#define FB_FMT_NRGBA64 3
#define FB_FMT_GRAY16 4
#define FB_FMT_CMYK 5
#define FB_FMT_PALETTED 6
FB_API int fb_convert_to_nrgba(int format, const uint8_t *pix, int stride, int w, int h, const uint16_t *palette16,
                        int ncolors, uint8_t *dst, int dstStride);

/* Reference-image session for the binary search of compress.go:45-74: `src` is uploaded and box-
 * downsampled ONCE (ssim.go:57 recomputes it every iteration); each iteration then sends only the
 * decoded candidate — as YCbCr planes (1.5 B/px at 4:2:0) or NRGBA — and gets SSIMFast(src, candidate)
 * (ssim.go:48-70) back.  Scores are identical to fb_ssim_fast on the converted image.  A handle is
 * bound to the device of the creating thread and may be used from one thread at a time. */
typedef struct fb_ssim_ref fb_ssim_ref;
FB_API int fb_ssim_ref_create(const uint8_t *src, int stride, int w, int h, fb_ssim_ref **out);
FB_API int fb_ssim_ref_score_ycbcr(const fb_ssim_ref *ref, const uint8_t *y, int yStride, const uint8_t *cb,
                            const uint8_t *cr, int cStride, int ratio, double *score);
FB_API int fb_ssim_ref_score_nrgba(const fb_ssim_ref *ref, const uint8_t *img, int stride, double *score);
FB_API void fb_ssim_ref_destroy(fb_ssim_ref *ref);

/* ---- Analyze — analyze.go:26-176 ---------------------------------------------- */

/* The fields of fennec.ImageStats. recommended_format / recommended_quality carry Go's numeric values
 * (types.go:36-42: JPEG = 1, PNG = 2; types.go:59-70: Balanced = 0, High = 3, Aggressive = 4). */
typedef struct fb_image_stats {
    int width, height;
    int has_alpha, is_grayscale, unique_colors;
    double entropy, edge_density, mean_brightness, contrast;
    int recommended_format, recommended_quality;
    double estimated_compression;
} fb_image_stats;

/* Analyze on a host NRGBA buffer. Integer fields and EdgeDensity are exact; MeanBrightness, Contrast and
 * Entropy agree with the reference's sequential float64 sums to ~1e-12 relative (summation order, log2). */
FB_API int fb_analyze(const uint8_t *pix, int stride, int w, int h, fb_image_stats *out);
/* Device-resident batch: writes n raw records (fb_analyze_raw_bytes() each) to device memory `raw`;
 * copy them to the host and turn each into ImageStats with fb_analyze_finish (host arithmetic only:
 * entropy from the histogram, sqrt, ratios, the recommendation rules of analyze.go:183-232). */
FB_API size_t fb_analyze_raw_bytes(void);
FB_API int fb_analyze_batch_dev(int device, void *stream, const uint8_t *imgs, int64_t imgStride, int rowStride,
                         int w, int h, int n, void *raw);
FB_API int fb_analyze_finish(const void *raw_host, int w, int h, fb_image_stats *out);

/* ---- ApplyOrientation — exif.go:176-203 (rotate / flip loops convert.go:186-256) ---- */

/* `orient` is the EXIF orientation value (exif.go:12-21: 2 FlipH, 3 Rotate180, 4 FlipV, 5 Transpose,
 * 6 Rotate90CW, 7 Transverse, 8 Rotate270CW). FB_IDENTITY for 1, 0 and unknown values (the reference returns
 * its input). dst must be dstW x dstH as reported by fb_orientation_dims (axes swap for 5-8). */
FB_API int fb_orientation_dims(int orient, int w, int h, int *dstW, int *dstH);
FB_API int fb_apply_orientation(const uint8_t *src, int srcStride, int w, int h, int orient, uint8_t *dst, int dstStride);
FB_API int fb_apply_orientation_batch_dev(int device, void *stream, const uint8_t *src, int64_t srcImgStride,
                                   int srcRowStride, int w, int h, int orient, uint8_t *dst, int64_t dstImgStride,
                                   int dstRowStride, int n);

/* ---- applyPalette + palettedToNRGBA — targetsize.go:479-545 ------------------------ */

/* `palette` holds ncolors (1..256) NRGBA entries with A == 255 (what medianCut returns, targetsize.go:400-413).
 * `indices` receives image.Paletted.Pix (first nearest entry by squared RGB distance, targetsize.go:499-510),
 * `dst` the palettedToNRGBA reconstruction; either may be NULL. */
FB_API int fb_apply_palette(const uint8_t *src, int srcStride, int w, int h, const uint8_t *palette, int ncolors,
                     uint8_t *indices, int idxStride, uint8_t *dst, int dstStride);
/* n device-resident images, one 256-entry palette slot (1024 bytes, device memory) per image.  Images of 2^17
 * pixels or more build per-palette cell lists (1 MiB each) in the calling thread's scratch arena —
 * fb_workspace_bytes("apply_palette", w, h, 0, 0, n); the result is identical either way. */
FB_API int fb_apply_palette_batch_dev(int device, void *stream, const uint8_t *src, int64_t imgStride, int rowStride,
                               int w, int h, int n, const uint8_t *palettes, int ncolors, uint8_t *indices,
                               int64_t idxImgStride, int idxRowStride, uint8_t *dst, int64_t dstImgStride,
                               int dstRowStride);

/* ---- device-resident batch entry points (configs 3-5 and the headline metric) ------------- */
/* n images (or pairs) of identical dims; image i starts at base + i*imgStride bytes. All pointers
 * are device pointers on `device`; `stream` is a cudaStream_t. Scores land in device memory. */

FB_API int fb_ssim_batch_dev(int device, void *stream, const uint8_t *a, const uint8_t *b,
                      int64_t imgStride, int rowStride, int w, int h, int n, double *scores);
FB_API int fb_ssim_fast_batch_dev(int device, void *stream, const uint8_t *a, const uint8_t *b,
                           int64_t imgStride, int rowStride, int w, int h, int n, double *scores);
FB_API int fb_msssim_batch_dev(int device, void *stream, const uint8_t *a, const uint8_t *b,
                        int64_t imgStride, int rowStride, int w, int h, int n, double *scores);
FB_API int fb_box_downsample_batch_dev(int device, void *stream, const uint8_t *src, int64_t srcImgStride,
                                int srcRowStride, int srcW, int srcH, uint8_t *dst, int64_t dstImgStride,
                                int dstRowStride, int dstW, int dstH, int n);
/* One MS-SSIM level step for both images of every pair (ssim.go:57-58 + 354-360): the SSIMFast thumbnail
 * (tw x th, from fb_ssim_fast_dims) AND the next level's image (w/2 x h/2) from a single read of the level
 * image. fb_msssim_batch_dev uses this internally; it is exported so the bytes can be checked directly. */
FB_API int fb_msssim_level_batch_dev(int device, void *stream, const uint8_t *a, const uint8_t *b, int64_t imgStride,
                              int rowStride, int w, int h, int n, uint8_t *thumbA, uint8_t *thumbB,
                              int64_t thumbImgStride, int thumbRowStride, int tw, int th, uint8_t *halfA,
                              uint8_t *halfB, int64_t halfImgStride, int halfRowStride);
/* Two level steps from ONE read of level l (box.cu: box_fused2_kernel): the <= 512 px thumbnails of level l and of
 * level l+1 (both tw x th) and the level-(l+2) image (w/4 x h/4), for both images of n pairs.  Level l+1 never exists
 * in memory; bytes equal boxDownsample applied step by step (ssim.go:57-58, 354-360).  Returns 1 with nothing written
 * when the geometry has no common box period (then call fb_msssim_level_batch_dev per level). */
FB_API int fb_msssim_level2_batch_dev(int device, void *stream, const uint8_t *a, const uint8_t *b, int64_t imgStride, int rowStride,
                               int w, int h, int n, uint8_t *thumb0A, uint8_t *thumb0B, uint8_t *thumb1A, uint8_t *thumb1B,
                               int64_t thumbImgStride, int thumbRowStride, int tw, int th, uint8_t *quarterA, uint8_t *quarterB,
                               int64_t quarterImgStride, int quarterRowStride);
/* convertToNRGBA for n device-resident YCbCr images (planes of image i at base + i*ImgStride). */
FB_API int fb_ycbcr_to_nrgba_batch_dev(int device, void *stream, const uint8_t *y, int64_t yImgStride, int yStride,
                                const uint8_t *cb, const uint8_t *cr, int64_t cImgStride, int cStride, int w,
                                int h, int ratio, uint8_t *dst, int64_t dstImgStride, int dstRowStride, int n);
/* n device-resident images of one format; Paletted: one 256-entry slot (2048 bytes, device memory) per image. */
FB_API int fb_convert_to_nrgba_batch_dev(int device, void *stream, int format, const uint8_t *pix, int64_t imgStride,
                                  int rowStride, int w, int h, int n, const uint16_t *palettes16, int ncolors,
                                  uint8_t *dst, int64_t dstImgStride, int dstRowStride);
FB_API int fb_gaussian_blur_batch_dev(int device, void *stream, const uint8_t *src, uint8_t *dst,
                               int64_t imgStride, int rowStride, int w, int h, int n,
                               const double *kernel_host, int radius);
FB_API int fb_sharpen_batch_dev(int device, void *stream, const uint8_t *src, uint8_t *dst,
                         int64_t imgStride, int rowStride, int w, int h, int n, double strength);
FB_API int fb_adaptive_sharpen_batch_dev(int device, void *stream, const uint8_t *src, uint8_t *dst,
                                  int64_t imgStride, int rowStride, int w, int h, int n, double strength);
FB_API int fb_lanczos_resize_batch_dev(int device, void *stream, const uint8_t *src, int64_t srcImgStride,
                                int srcRowStride, int srcW, int srcH, uint8_t *dst, int64_t dstImgStride,
                                int dstRowStride, int dstW, int dstH, int n);
/* Bytes of scratch the _dev call above needs for these dims. The library keeps ONE grow-only scratch arena per
 * calling thread and device and re-uses it from the start in every call: work enqueued by one thread must
 * therefore go to one stream at a time (or be synchronised between streams); different threads never share. */
FB_API size_t fb_workspace_bytes(const char *op, int w, int h, int dstW, int dstH, int n);

/* ---- batch sharder (batch.go:58-128) ------------------------------------------------------ */

/* Static contiguous partition of n_items over n_shards (GPUs / ranks): shard s owns
 * [begin, end). Keeps input order, so results[idx] semantics (batch.go:71,108) hold after a
 * concatenating gather. */
FB_API int fb_batch_shard(int n_items, int n_shards, int shard, int *begin, int *end);

/* ---- host-buffer batches: CompressBatch's worker pool behind one call (batch.go:58-128) ---- */

/* One call shards `n` items over every initialised device (fb_batch_shard: contiguous blocks, input order kept) and
 * runs them on the library's own worker threads (workers_per_device per GPU, default 4; they pull indices from a
 * shared counter like the reference's channel of indices, batch.go:72-81).  Results are written at their input
 * index (batch.go:108-113).  status[i] (optional) receives the item's return code; a failing item does not stop
 * the others (batch.go:107-113); if *cancel becomes non-zero, items not yet started get FB_E_CANCELLED
 * (batch.go:90-98); on_item(completed, total, user) fires after every item and may run concurrently on several
 * worker threads (batch.go:115-121).  Returns the number of failed / cancelled items (0 = all good) or a negative
 * error for bad arguments.  This is what a cgo CompressBatch binds when it wants all GPUs of the box from ONE
 * process; the multi-process form (one rank per GPU) uses fb_batch_shard + an all-gather instead. */
typedef void (*fb_progress_fn)(int completed, int total, void *user);
typedef struct fb_batch_opts {
    int workers_per_device;       /* <= 0: 4 */
    const volatile int *cancel;   /* optional cancellation flag (ctx.Done) */
    fb_progress_fn on_item;       /* optional progress callback (BatchOptions.OnItem, batch.go:40) */
    void *user;
} fb_batch_opts;

typedef struct fb_pair {          /* two NRGBA images of equal dims */
    const uint8_t *a; int strideA;
    const uint8_t *b; int strideB;
    int w, h;
} fb_pair;
#define FB_OP_SSIM 0      /* fennec.SSIM (ssim.go:24) per pair */
#define FB_OP_SSIM_FAST 1 /* fennec.SSIMFast (ssim.go:48) */
#define FB_OP_MSSSIM 2    /* fennec.MSSSIM (ssim.go:313) */
FB_API int fb_score_batch_host(int op, const fb_pair *pairs, int n, double *scores, int *status, const fb_batch_opts *opts);

typedef struct fb_resize_item {   /* lanczosResize (resize.go:37) of one image; dst allocated by the caller */
    const uint8_t *src; int srcStride, srcW, srcH;
    uint8_t *dst; int dstStride, dstW, dstH;
} fb_resize_item;
FB_API int fb_lanczos_resize_batch_host(const fb_resize_item *items, int n, int *status, const fb_batch_opts *opts);

typedef struct fb_effect_item {   /* src -> dst, same dims */
    const uint8_t *src; int srcStride;
    uint8_t *dst; int dstStride;
    int w, h;
} fb_effect_item;
#define FB_FX_GAUSSIAN_BLUR 0     /* param = sigma  (effects.go:146) */
#define FB_FX_SHARPEN 1           /* param = strength (effects.go:10) */
#define FB_FX_ADAPTIVE_SHARPEN 2  /* param = strength (effects.go:49) */
/* status[i] may be FB_IDENTITY: the reference returns its input (dst not written). */
FB_API int fb_effect_batch_host(int effect, double param, const fb_effect_item *items, int n, int *status, const fb_batch_opts *opts);

/* Pinned (page-locked) host memory: uploads from it are DMA'ed directly; uploads from ordinary pageable memory
 * (a Go Pix slice) are staged through the library's own pinned chunks instead.  A Go caller can wrap the pointer
 * with unsafe.Slice and decode into it. */
FB_API void *fb_alloc_pinned(size_t bytes);
FB_API void fb_free_pinned(void *p);

/* Diagnostics: idle per-thread contexts (stream + arenas) parked in the pool, and cached Lanczos table pairs. */
FB_API int fb_debug_pool_size(void);
FB_API int fb_debug_table_count(void);

/* Kernel launches issued by this thread since the last call (bench.py's gpu_launches). */
FB_API long long fb_take_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
