/*
 * fennec_oracle.h — CPU restatement of the fennec hot path (TEST INFRASTRUCTURE ONLY).
 *
 * PARITY UNPINNED: the reference (shamspias/fennec @ 98234f2c) is pure Go and no Go
 * toolchain exists in the build image, so this oracle could not be checked against
 * outputs of the reference itself; the reference's own tests hold no golden vectors
 * for this path (only inequalities).  It is pinned instead by (i) a
 * second, independent NumPy restatement (oracle/np_restatement.py) that must agree
 * bit-for-bit on pixels and to <=1e-12 on scores, (ii) the reference's inequality
 * tests carried over, (iii) analytic known answers (e.g. black-vs-white SSIM).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call this library.  The product (libfennec_b200.so) never does.
 *
 * Arithmetic rules: IEEE binary64, round-to-nearest-even, no FMA
 * contraction (-ffp-contract=off == Go on amd64), source-order evaluation, ascending
 * tap order, C cast for Go's int(float64), round() for math.Round (half away from 0).
 */
#ifndef FENNEC_ORACLE_H
#define FENNEC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Worker count that stands in for runtime.GOMAXPROCS(0) (resize.go:206, ssim.go:84). */
void fo_set_procs(int procs);
int fo_get_procs(void);

/* convert.go:149-158 */
uint8_t fo_clampf(double x);

/* ssim.go:223-241 — size*size weights, row-major, normalised by their running sum. */
void fo_gaussian_kernel(int size, double sigma, double *out);

/* ssim.go:207-220 — lum[y*w+x], honours stride. */
void fo_to_luminance(const uint8_t *pix, int stride, int w, int h, double *lum);

/* ssim.go:73-166 — procs<=0 means fo_get_procs(). */
double fo_windowed_ssim(const double *lumA, const double *lumB, int w, int h, int procs);

/* ssim.go:169-204 — flat iteration over Pix (compact images: stride == 4*w). */
double fo_pixel_ssim(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h);

/* ssim.go:24-43 for equal-sized NRGBA inputs (the Lanczos pre-resize at :31-33 is done by the caller). */
double fo_ssim(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h);

/* ssim.go:48-70 */
double fo_ssim_fast(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h);

/* ssim.go:52-56 — target dims of SSIMFast's downsample; returns 1 if a downsample happens. */
int fo_ssim_fast_dims(int w, int h, int *newW, int *newH);

/* ssim.go:244-309 — dst must hold dstH rows of dstStride bytes, pre-zeroed by the caller
 * (Go's image.NewNRGBA zero-fills).  Returns 0, or 1 when any dim <= 0 (empty image). */
int fo_box_downsample(const uint8_t *src, int srcStride, int srcW, int srcH,
                      uint8_t *dst, int dstStride, int dstW, int dstH);

/* ssim.go:313-365 for equal-sized inputs. */
double fo_msssim(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h);

/* effects.go:153-165 — radius = ceil(3*sigma); kernel has 2*radius+1 entries. */
int fo_blur_radius(double sigma);
void fo_blur_kernel(double sigma, int radius, double *kernel);

/* effects.go:146-220 with the kernel supplied by the caller (fo_blur_kernel or a Go-built table). */
void fo_gaussian_blur_k(const uint8_t *src, int srcStride, int w, int h,
                        const double *kernel, int radius, uint8_t *dst, int dstStride);
/* effects.go:146-220; returns 1 when sigma <= 0 (reference returns the SAME pointer, dst untouched). */
int fo_gaussian_blur(const uint8_t *src, int srcStride, int w, int h, double sigma,
                     uint8_t *dst, int dstStride);

/* effects.go:116-141 */
void fo_blur3x3(const uint8_t *src, int srcStride, int w, int h, uint8_t *dst, int dstStride);

/* effects.go:10-45 / 49-90; return 1 for the identity guards (strength<=0, w<3, h<3), dst untouched. */
int fo_sharpen(const uint8_t *src, int srcStride, int w, int h, double strength,
               uint8_t *dst, int dstStride);
int fo_adaptive_sharpen(const uint8_t *src, int srcStride, int w, int h, double strength,
                        uint8_t *dst, int dstStride);

/* resize.go:57-69 */
double fo_lanczos_kernel(double x);

/* resize.go:164-197 as CSR: for d in [0,dstSize): taps index[start[d] .. start[d+1]) with weights.
 * start has dstSize+1 entries; index/weight must hold fo_lanczos_weights_cap(dstSize,srcSize) entries.
 * Returns the number of entries written. */
int fo_lanczos_weights_cap(int dstSize, int srcSize);
int fo_lanczos_weights(int dstSize, int srcSize, int *start, int *index, double *weight);

/* resize.go:77-118 / 121-161 — dst pre-zeroed by the caller (pixels with a <= 0.5 stay 0). */
void fo_resize_h(const uint8_t *src, int srcStride, int srcW, int srcH,
                 uint8_t *dst, int dstStride, int dstW);
void fo_resize_v(const uint8_t *src, int srcStride, int srcW, int srcH,
                 uint8_t *dst, int dstStride, int dstH);
/* resize.go:37-53 — returns 1 for the empty-image guard (any dim <= 0). dst pre-zeroed. */
int fo_lanczos_resize(const uint8_t *src, int srcStride, int srcW, int srcH,
                      uint8_t *dst, int dstStride, int dstW, int dstH);

/* resize.go:12-32 — writes the dims smartResize would produce; returns 1 if it is a no-op. */
int fo_smart_resize_dims(int srcW, int srcH, int maxW, int maxH, int *dstW, int *dstH);

/* convertToNRGBA (convert.go:34-64) on what jpeg.Decode returns.  ratio = Go's
 * image.YCbCrSubsampleRatio constant (0: 4:4:4, 1: 4:2:2, 2: 4:2:0, 3: 4:4:0, 4: 4:1:1, 5: 4:1:0).  Go stdlib arithmetic
 * (image/color/ycbcr.go, Go 1.25.5; not vendored) restated from the published source — see the .c file. */
int fo_ycbcr_to_nrgba(const uint8_t *y, int yStride, const uint8_t *cb, const uint8_t *cr, int cStride,
                      int w, int h, int ratio, uint8_t *dst, int dstStride);
void fo_gray_to_nrgba(const uint8_t *g, int gStride, int w, int h, uint8_t *dst, int dstStride);
/* convertToNRGBA for *image.RGBA (1), RGBA64 (2), NRGBA64 (3), Gray16 (4), CMYK (5), Paletted (6; pal16 = ncolors x
 * 4 uint16 = Palette[i].RGBA()).  0, -1 bad arguments, -2 palette index out of range (Go panics). */
int fo_convert_to_nrgba(int fmt, const uint8_t *pix, int stride, int w, int h, const uint16_t *pal16, int ncolors,
                        uint8_t *dst, int dstStride);

/* Analyze (analyze.go:26-176) and the recommendation rules (analyze.go:183-232). */
typedef struct {
    int width, height;
    int has_alpha, is_grayscale, unique_colors;
    double entropy, edge_density, mean_brightness, contrast;
    int recommended_format;   /* Go's Format value: 1 = JPEG, 2 = PNG (types.go:36-42) */
    int recommended_quality;  /* Go's Quality value: 0 = Balanced, 3 = High, 4 = Aggressive (types.go:59-70) */
    double estimated_compression;
    double histogram[256];    /* luminance histogram, bins int(lum + 0.5) */
} fo_image_stats;
void fo_analyze(const uint8_t *pix, int stride, int w, int h, fo_image_stats *st);
void fo_recommend(fo_image_stats *st);

/* ApplyOrientation (exif.go:176-203); returns 1 for the identity orientations. */
int fo_apply_orientation(const uint8_t *src, int srcStride, int w, int h, int orient, uint8_t *dst, int dstStride);

/* applyPalette + palettedToNRGBA (targetsize.go:479-545). */
void fo_apply_palette(const uint8_t *src, int srcStride, int w, int h, const uint8_t *palette, int ncolors,
                      uint8_t *idx, int idxStride, uint8_t *out, int outStride);

#ifdef __cplusplus
}
#endif
#endif
