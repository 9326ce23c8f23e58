/*
 * fennec_oracle.c — CPU restatement of the fennec hot path in plain C (TEST INFRASTRUCTURE ONLY).
 * See fennec_oracle.h for the rules and the "parity unpinned" statement.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared -pthread (oracle/Makefile).
 * Every function cites the reference file:line (under /root/reference) it restates.
 */
#include "fennec_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * parallelDo — resize.go:200-239: contiguous index blocks over GOMAXPROCS workers.
 * Every index writes disjoint output, so the partition does not change any result.
 * ---------------------------------------------------------------------------------------- */
static int g_procs = 8;

void fo_set_procs(int procs) { g_procs = procs < 1 ? 1 : procs; }
int fo_get_procs(void) { return g_procs; }

typedef void (*fo_index_fn)(int i, void *ctx);

typedef struct {
    int from, to;
    fo_index_fn fn;
    void *ctx;
} fo_span;

static void *fo_span_run(void *arg) {
    fo_span *s = (fo_span *)arg;
    for (int i = s->from; i < s->to; i++) s->fn(i, s->ctx);
    return NULL;
}

static void parallel_do(int start, int stop, fo_index_fn fn, void *ctx) {
    int count = stop - start;
    if (count <= 0) return;
    int procs = g_procs;
    if (procs > count) procs = count;
    if (procs <= 1) {
        for (int i = start; i < stop; i++) fn(i, ctx);
        return;
    }
    int batch = (count + procs - 1) / procs;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)procs);
    fo_span *sp = (fo_span *)malloc(sizeof(fo_span) * (size_t)procs);
    int launched = 0;
    for (int p = 0; p < procs; p++) {
        int b0 = start + p * batch;
        int b1 = b0 + batch;
        if (b1 > stop) b1 = stop;
        if (b0 >= b1) continue;
        sp[launched].from = b0;
        sp[launched].to = b1;
        sp[launched].fn = fn;
        sp[launched].ctx = ctx;
        if (pthread_create(&tid[launched], NULL, fo_span_run, &sp[launched]) != 0) {
            fo_span_run(&sp[launched]); /* degrade to inline execution */
            sp[launched].from = sp[launched].to;
            tid[launched] = pthread_self();
        }
        launched++;
    }
    for (int p = 0; p < launched; p++)
        if (!pthread_equal(tid[p], pthread_self())) pthread_join(tid[p], NULL);
    free(tid);
    free(sp);
}

/* convert.go:149-158 — int64(math.Round(x)) clamped to [0,255]; Round is half away from zero. */
uint8_t fo_clampf(double x) {
    long long v = (long long)round(x);
    if (v > 255) return 255;
    if (v < 0) return 0;
    return (uint8_t)v;
}

/* ssim.go:10-17 — Go evaluates the untyped constant expressions exactly, then rounds once. */
static const double kC1 = 6.5025;
static const double kC2 = 58.5225;

/* ssim.go:223-241 */
void fo_gaussian_kernel(int size, double sigma, double *out) {
    int half = size / 2;
    double sum = 0.0;
    int idx = 0;
    for (int y = -half; y < half; y++) {
        for (int x = -half; x < half; x++) {
            double val = exp(-(double)(x * x + y * y) / (2 * sigma * sigma));
            out[idx] = val;
            sum += val;
            idx++;
        }
    }
    for (int i = 0; i < size * size; i++) out[i] /= sum;
}

/* The BT.601 expression used at ssim.go:179-180,216 and effects.go:96: (0.299*R + 0.587*G) + 0.114*B. */
static inline double luma_of(const uint8_t *p) {
    return 0.299 * (double)p[0] + 0.587 * (double)p[1] + 0.114 * (double)p[2];
}

/* ssim.go:207-220 */
void fo_to_luminance(const uint8_t *pix, int stride, int w, int h, double *lum) {
    for (int y = 0; y < h; y++) {
        const uint8_t *row = pix + (size_t)y * (size_t)stride;
        for (int x = 0; x < w; x++) lum[(size_t)y * w + x] = luma_of(row + x * 4);
    }
}

/* ssim.go:73-166 */
typedef struct {
    const double *lumA, *lumB, *kernel;
    int w, h, half, rowsPerProc;
    double sum;
    long long count;
    int proc;
} fo_ssim_job;

static void *fo_ssim_worker(void *arg) {
    fo_ssim_job *j = (fo_ssim_job *)arg;
    const int w = j->w, h = j->h, half = j->half;
    int startY = half + j->proc * j->rowsPerProc; /* ssim.go:101-105 */
    int endY = startY + j->rowsPerProc;
    if (endY > h - half) endY = h - half;
    double localSum = 0.0;
    long long localCount = 0;
    for (int y = startY; y < endY; y++) {
        for (int x = half; x < w - half; x++) {
            double muA = 0.0, muB = 0.0;
            double sigAA = 0.0, sigBB = 0.0, sigAB = 0.0;
            int ki = 0;
            for (int wy = -half; wy < half; wy++) { /* ssim.go:115-126 */
                for (int wx = -half; wx < half; wx++) {
                    size_t idx = (size_t)(y + wy) * w + (x + wx);
                    double weight = j->kernel[ki];
                    muA += j->lumA[idx] * weight;
                    muB += j->lumB[idx] * weight;
                    ki++;
                }
            }
            ki = 0;
            for (int wy = -half; wy < half; wy++) { /* ssim.go:128-140 */
                for (int wx = -half; wx < half; wx++) {
                    size_t idx = (size_t)(y + wy) * w + (x + wx);
                    double weight = j->kernel[ki];
                    double da = j->lumA[idx] - muA;
                    double db = j->lumB[idx] - muB;
                    sigAA += da * da * weight;
                    sigBB += db * db * weight;
                    sigAB += da * db * weight;
                    ki++;
                }
            }
            double num = (2 * muA * muB + kC1) * (2 * sigAB + kC2); /* ssim.go:142-145 */
            double den = (muA * muA + muB * muB + kC1) * (sigAA + sigBB + kC2);
            localSum += num / den;
            localCount++;
        }
    }
    j->sum = localSum;
    j->count = localCount;
    return NULL;
}

double fo_windowed_ssim(const double *lumA, const double *lumB, int w, int h, int procs) {
    const int windowSize = 8;
    const int half = windowSize / 2;
    double kernel[64];
    fo_gaussian_kernel(windowSize, 1.5, kernel);

    if (procs <= 0) procs = g_procs;
    int rows = h - windowSize + 1; /* ssim.go:85-91 */
    if (procs > rows) procs = rows;
    if (procs < 1) procs = 1;
    int rowsPerProc = (rows + procs - 1) / procs;

    fo_ssim_job *jobs = (fo_ssim_job *)calloc((size_t)procs, sizeof(fo_ssim_job));
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)procs);
    for (int p = 0; p < procs; p++) {
        jobs[p].lumA = lumA;
        jobs[p].lumB = lumB;
        jobs[p].kernel = kernel;
        jobs[p].w = w;
        jobs[p].h = h;
        jobs[p].half = half;
        jobs[p].rowsPerProc = rowsPerProc;
        jobs[p].proc = p;
    }
    if (procs == 1) {
        fo_ssim_worker(&jobs[0]);
    } else {
        for (int p = 0; p < procs; p++)
            if (pthread_create(&tid[p], NULL, fo_ssim_worker, &jobs[p]) != 0) {
                fo_ssim_worker(&jobs[p]);
                tid[p] = pthread_self();
            }
        for (int p = 0; p < procs; p++)
            if (!pthread_equal(tid[p], pthread_self())) pthread_join(tid[p], NULL);
    }
    double totalSum = 0.0; /* ssim.go:155-165 */
    long long totalCount = 0;
    for (int p = 0; p < procs; p++) {
        totalSum += jobs[p].sum;
        totalCount += jobs[p].count;
    }
    free(jobs);
    free(tid);
    if (totalCount == 0) return 1.0;
    return totalSum / (double)totalCount;
}

/* ssim.go:169-204 — the reference walks Pix as a flat array (i += 4), ignoring Stride; for
 * compact images (Stride == 4*w, the only kind the hot path produces) that is every pixel once.
 * For padded strides we walk rows, which is what the flat walk means for a compact copy. */
double fo_pixel_ssim(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h) {
    double n = (double)(w * h);
    if (n == 0) return 1.0;
    double muA = 0.0, muB = 0.0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            double la = luma_of(a + (size_t)y * strideA + x * 4);
            double lb = luma_of(b + (size_t)y * strideB + x * 4);
            muA += la;
            muB += lb;
        }
    muA /= n;
    muB /= n;
    double sigAA = 0.0, sigBB = 0.0, sigAB = 0.0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            double la = luma_of(a + (size_t)y * strideA + x * 4);
            double lb = luma_of(b + (size_t)y * strideB + x * 4);
            double da = la - muA;
            double db = lb - muB;
            sigAA += da * da;
            sigBB += db * db;
            sigAB += da * db;
        }
    sigAA /= n;
    sigBB /= n;
    sigAB /= n;
    double num = (2 * muA * muB + kC1) * (2 * sigAB + kC2);
    double den = (muA * muA + muB * muB + kC1) * (sigAA + sigBB + kC2);
    return num / den;
}

/* ssim.go:35-42 / 62-69 — shared tail of SSIM and SSIMFast. */
static double ssim_tail(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h) {
    if (w < 8 || h < 8) return fo_pixel_ssim(a, strideA, b, strideB, w, h);
    double *lumA = (double *)malloc(sizeof(double) * (size_t)w * h);
    double *lumB = (double *)malloc(sizeof(double) * (size_t)w * h);
    fo_to_luminance(a, strideA, w, h, lumA);
    fo_to_luminance(b, strideB, w, h, lumB);
    double r = fo_windowed_ssim(lumA, lumB, w, h, 0);
    free(lumA);
    free(lumB);
    return r;
}

/* ssim.go:24-43 (equal dims) */
double fo_ssim(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h) {
    return ssim_tail(a, strideA, b, strideB, w, h);
}

/* ssim.go:52-56 */
int fo_ssim_fast_dims(int w, int h, int *newW, int *newH) {
    const int maxDim = 512;
    if (w > maxDim || h > maxDim) {
        double scale = (double)maxDim / fmax((double)w, (double)h);
        *newW = (int)fmax(8, round((double)w * scale));
        *newH = (int)fmax(8, round((double)h * scale));
        return 1;
    }
    *newW = w;
    *newH = h;
    return 0;
}

/* ssim.go:286-309 */
static void average_box_pixel(const uint8_t *src, int srcStride, uint8_t *dst, int dstStride,
                              int dx, int dy, int sx0, int sx1, int sy0, int sy1) {
    double rSum = 0, gSum = 0, bSum = 0, aSum = 0, count = 0;
    for (int sy = sy0; sy < sy1; sy++)
        for (int sx = sx0; sx < sx1; sx++) {
            const uint8_t *p = src + (size_t)sy * srcStride + sx * 4;
            rSum += (double)p[0];
            gSum += (double)p[1];
            bSum += (double)p[2];
            aSum += (double)p[3];
            count++;
        }
    if (count > 0) {
        double inv = 1.0 / count;
        uint8_t *q = dst + (size_t)dy * dstStride + dx * 4;
        q[0] = fo_clampf(rSum * inv);
        q[1] = fo_clampf(gSum * inv);
        q[2] = fo_clampf(bSum * inv);
        q[3] = fo_clampf(aSum * inv);
    }
}

/* ssim.go:244-284 (serial in the reference) */
int fo_box_downsample(const uint8_t *src, int srcStride, int srcW, int srcH,
                      uint8_t *dst, int dstStride, int dstW, int dstH) {
    if (srcW <= 0 || srcH <= 0 || dstW <= 0 || dstH <= 0) return 1;
    double xRatio = (double)srcW / (double)dstW;
    double yRatio = (double)srcH / (double)dstH;
    for (int dy = 0; dy < dstH; dy++) {
        int sy0 = (int)((double)dy * yRatio);
        int sy1 = (int)((double)(dy + 1) * yRatio);
        if (sy1 > srcH) sy1 = srcH;
        if (sy0 >= sy1) sy0 = sy1 - 1;
        if (sy0 < 0) sy0 = 0;
        for (int dx = 0; dx < dstW; dx++) {
            int sx0 = (int)((double)dx * xRatio);
            int sx1 = (int)((double)(dx + 1) * xRatio);
            if (sx1 > srcW) sx1 = srcW;
            if (sx0 >= sx1) sx0 = sx1 - 1;
            if (sx0 < 0) sx0 = 0;
            average_box_pixel(src, srcStride, dst, dstStride, dx, dy, sx0, sx1, sy0, sy1);
        }
    }
    return 0;
}

/* ssim.go:48-70 */
double fo_ssim_fast(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h) {
    int nw, nh;
    if (fo_ssim_fast_dims(w, h, &nw, &nh)) {
        uint8_t *da = (uint8_t *)calloc((size_t)nw * nh, 4);
        uint8_t *db = (uint8_t *)calloc((size_t)nw * nh, 4);
        fo_box_downsample(a, strideA, w, h, da, nw * 4, nw, nh);
        fo_box_downsample(b, strideB, w, h, db, nw * 4, nw, nh);
        double r = ssim_tail(da, nw * 4, db, nw * 4, nw, nh);
        free(da);
        free(db);
        return r;
    }
    return ssim_tail(a, strideA, b, strideB, w, h);
}

/* ssim.go:313-365 (equal dims) */
double fo_msssim(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w0, int h0) {
    double weights[5] = {0.0448, 0.2856, 0.3001, 0.2363, 0.1333};
    int levels = 5, nweights = 5;
    int w = w0, h = h0;
    for (int i = 0; i < levels - 1; i++) { /* ssim.go:327-342 */
        int minDim = (int)fmin((double)w, (double)h);
        if (minDim < 8) {
            nweights = i + 1;
            double sum = 0.0;
            for (int j = 0; j < nweights; j++) sum += weights[j];
            for (int j = 0; j < nweights; j++) weights[j] /= sum;
            break;
        }
        w /= 2;
        h /= 2;
    }
    /* ssim.go:345-346 — mutable compact copies */
    int cw = w0, ch = h0;
    uint8_t *ca = (uint8_t *)malloc((size_t)cw * ch * 4 + 4);
    uint8_t *cb = (uint8_t *)malloc((size_t)cw * ch * 4 + 4);
    for (int y = 0; y < ch; y++) {
        memcpy(ca + (size_t)y * cw * 4, a + (size_t)y * strideA, (size_t)cw * 4);
        memcpy(cb + (size_t)y * cw * 4, b + (size_t)y * strideB, (size_t)cw * 4);
    }
    double result = 0.0;
    for (int i = 0; i < nweights; i++) { /* ssim.go:349-362 */
        double s = fo_ssim_fast(ca, cw * 4, cb, cw * 4, cw, ch);
        result += weights[i] * log(fmax(s, 1e-10));
        if (i < nweights - 1) {
            int nw = cw / 2, nh = ch / 2;
            if (nw < 8 || nh < 8) break;
            uint8_t *na = (uint8_t *)calloc((size_t)nw * nh, 4);
            uint8_t *nb = (uint8_t *)calloc((size_t)nw * nh, 4);
            fo_box_downsample(ca, cw * 4, cw, ch, na, nw * 4, nw, nh);
            fo_box_downsample(cb, cw * 4, cw, ch, nb, nw * 4, nw, nh);
            free(ca);
            free(cb);
            ca = na;
            cb = nb;
            cw = nw;
            ch = nh;
        }
    }
    free(ca);
    free(cb);
    return exp(result);
}

/* ------------------------------------------------------------------------------------------
 * effects.go
 * ---------------------------------------------------------------------------------------- */

/* effects.go:153 */
int fo_blur_radius(double sigma) { return (int)ceil(sigma * 3); }

/* effects.go:155-165 */
void fo_blur_kernel(double sigma, int radius, double *kernel) {
    int kernelSize = radius * 2 + 1;
    double sum = 0.0;
    for (int i = 0; i < kernelSize; i++) {
        double x = (double)(i - radius);
        kernel[i] = exp(-(x * x) / (2 * sigma * sigma));
        sum += kernel[i];
    }
    for (int i = 0; i < kernelSize; i++) kernel[i] /= sum;
}

typedef struct {
    const uint8_t *src; /* image the taps read */
    int srcStride;
    const uint8_t *alpha; /* image alpha is copied from (always the ORIGINAL source) */
    int alphaStride;
    uint8_t *dst;
    int dstStride;
    int w, h, radius;
    const double *kernel;
} fo_blur_ctx;

/* effects.go:169-191 — one row of the horizontal pass */
static void blur_h_row(int y, void *vctx) {
    fo_blur_ctx *c = (fo_blur_ctx *)vctx;
    int kernelSize = c->radius * 2 + 1;
    for (int x = 0; x < c->w; x++) {
        double r = 0, g = 0, b = 0;
        for (int k = 0; k < kernelSize; k++) {
            int sx = x + k - c->radius;
            if (sx < 0) sx = 0;
            else if (sx >= c->w) sx = c->w - 1;
            const uint8_t *p = c->src + (size_t)y * c->srcStride + sx * 4;
            double wt = c->kernel[k];
            r += (double)p[0] * wt;
            g += (double)p[1] * wt;
            b += (double)p[2] * wt;
        }
        uint8_t *q = c->dst + (size_t)y * c->dstStride + x * 4;
        q[0] = fo_clampf(r);
        q[1] = fo_clampf(g);
        q[2] = fo_clampf(b);
        q[3] = c->alpha[(size_t)y * c->alphaStride + x * 4 + 3];
    }
}

/* effects.go:195-217 — one column of the vertical pass */
static void blur_v_col(int x, void *vctx) {
    fo_blur_ctx *c = (fo_blur_ctx *)vctx;
    int kernelSize = c->radius * 2 + 1;
    for (int y = 0; y < c->h; y++) {
        double r = 0, g = 0, b = 0;
        for (int k = 0; k < kernelSize; k++) {
            int sy = y + k - c->radius;
            if (sy < 0) sy = 0;
            else if (sy >= c->h) sy = c->h - 1;
            const uint8_t *p = c->src + (size_t)sy * c->srcStride + x * 4;
            double wt = c->kernel[k];
            r += (double)p[0] * wt;
            g += (double)p[1] * wt;
            b += (double)p[2] * wt;
        }
        uint8_t *q = c->dst + (size_t)y * c->dstStride + x * 4;
        q[0] = fo_clampf(r);
        q[1] = fo_clampf(g);
        q[2] = fo_clampf(b);
        q[3] = c->alpha[(size_t)y * c->alphaStride + x * 4 + 3];
    }
}

void fo_gaussian_blur_k(const uint8_t *src, int srcStride, int w, int h,
                        const double *kernel, int radius, uint8_t *dst, int dstStride) {
    if (w <= 0 || h <= 0) return;
    uint8_t *tmp = (uint8_t *)calloc((size_t)w * h, 4);
    fo_blur_ctx c;
    c.src = src; c.srcStride = srcStride; c.alpha = src; c.alphaStride = srcStride;
    c.dst = tmp; c.dstStride = w * 4; c.w = w; c.h = h; c.radius = radius; c.kernel = kernel;
    parallel_do(0, h, blur_h_row, &c);
    c.src = tmp; c.srcStride = w * 4;
    c.dst = dst; c.dstStride = dstStride;
    parallel_do(0, w, blur_v_col, &c);
    free(tmp);
}

int fo_gaussian_blur(const uint8_t *src, int srcStride, int w, int h, double sigma,
                     uint8_t *dst, int dstStride) {
    if (sigma <= 0) return 1; /* effects.go:147-149 — same pointer */
    int radius = fo_blur_radius(sigma);
    double *kernel = (double *)malloc(sizeof(double) * (size_t)(2 * radius + 1));
    fo_blur_kernel(sigma, radius, kernel);
    fo_gaussian_blur_k(src, srcStride, w, h, kernel, radius, dst, dstStride);
    free(kernel);
    return 0;
}

typedef struct {
    const uint8_t *src;
    int srcStride;
    const uint8_t *blur;
    int blurStride;
    uint8_t *dst;
    int dstStride;
    int w, h;
    double amount;
} fo_fx_ctx;

/* effects.go:122-139 — one interior row */
static void blur3_row(int y, void *vctx) {
    fo_fx_ctx *c = (fo_fx_ctx *)vctx;
    const uint8_t *s = c->src;
    int st = c->srcStride;
    for (int x = 1; x < c->w - 1; x++)
        for (int ch = 0; ch < 3; ch++) {
            double sum = 0;
            sum += (double)s[(size_t)(y - 1) * st + (x - 1) * 4 + ch] * 1;
            sum += (double)s[(size_t)(y - 1) * st + (x)*4 + ch] * 2;
            sum += (double)s[(size_t)(y - 1) * st + (x + 1) * 4 + ch] * 1;
            sum += (double)s[(size_t)(y)*st + (x - 1) * 4 + ch] * 2;
            sum += (double)s[(size_t)(y)*st + (x)*4 + ch] * 4;
            sum += (double)s[(size_t)(y)*st + (x + 1) * 4 + ch] * 2;
            sum += (double)s[(size_t)(y + 1) * st + (x - 1) * 4 + ch] * 1;
            sum += (double)s[(size_t)(y + 1) * st + (x)*4 + ch] * 2;
            sum += (double)s[(size_t)(y + 1) * st + (x + 1) * 4 + ch] * 1;
            c->dst[(size_t)y * c->dstStride + x * 4 + ch] = fo_clampf(sum / 16.0);
        }
}

/* effects.go:116-141 */
void fo_blur3x3(const uint8_t *src, int srcStride, int w, int h, uint8_t *dst, int dstStride) {
    for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * dstStride, src + (size_t)y * srcStride, (size_t)w * 4);
    fo_fx_ctx c;
    memset(&c, 0, sizeof c);
    c.src = src; c.srcStride = srcStride; c.dst = dst; c.dstStride = dstStride; c.w = w; c.h = h;
    parallel_do(1, h - 1, blur3_row, &c);
}

/* effects.go:28-42 */
static void sharpen_row(int y, void *vctx) {
    fo_fx_ctx *c = (fo_fx_ctx *)vctx;
    for (int x = 0; x < c->w; x++) {
        const uint8_t *s = c->src + (size_t)y * c->srcStride + x * 4;
        const uint8_t *bl = c->blur + (size_t)y * c->blurStride + x * 4;
        uint8_t *d = c->dst + (size_t)y * c->dstStride + x * 4;
        for (int ch = 0; ch < 3; ch++) {
            double orig = (double)s[ch];
            double blur = (double)bl[ch];
            double val = orig + c->amount * (orig - blur);
            d[ch] = fo_clampf(val);
        }
        d[3] = s[3];
    }
}

/* effects.go:10-45 */
int fo_sharpen(const uint8_t *src, int srcStride, int w, int h, double strength,
               uint8_t *dst, int dstStride) {
    if (strength <= 0) return 1;
    if (strength > 1) strength = 1;
    if (w < 3 || h < 3) return 1;
    uint8_t *blurred = (uint8_t *)malloc((size_t)w * h * 4);
    fo_blur3x3(src, srcStride, w, h, blurred, w * 4);
    fo_fx_ctx c;
    c.src = src; c.srcStride = srcStride; c.blur = blurred; c.blurStride = w * 4;
    c.dst = dst; c.dstStride = dstStride; c.w = w; c.h = h;
    c.amount = 1.0 + strength * 1.5;
    parallel_do(0, h, sharpen_row, &c);
    free(blurred);
    return 0;
}

/* effects.go:93-112 */
static double local_edge_strength(const uint8_t *pix, int stride, int x, int y) {
#define FO_LUM(px, py) luma_of(pix + (size_t)(py)*stride + (px)*4)
    double gx = -FO_LUM(x - 1, y - 1) + FO_LUM(x + 1, y - 1) -
                2 * FO_LUM(x - 1, y) + 2 * FO_LUM(x + 1, y) -
                FO_LUM(x - 1, y + 1) + FO_LUM(x + 1, y + 1);
    double gy = -FO_LUM(x - 1, y - 1) - 2 * FO_LUM(x, y - 1) - FO_LUM(x + 1, y - 1) +
                FO_LUM(x - 1, y + 1) + 2 * FO_LUM(x, y + 1) + FO_LUM(x + 1, y + 1);
#undef FO_LUM
    double mag = sqrt(gx * gx + gy * gy);
    double normalized = mag / 400.0;
    if (normalized > 1) normalized = 1;
    return normalized;
}

/* effects.go:70-87 */
static void adaptive_row(int y, void *vctx) {
    fo_fx_ctx *c = (fo_fx_ctx *)vctx;
    for (int x = 1; x < c->w - 1; x++) {
        const uint8_t *s = c->src + (size_t)y * c->srcStride + x * 4;
        double edgeStr = local_edge_strength(c->src, c->srcStride, x, y);
        double localAmount = c->amount * edgeStr;
        const uint8_t *bl = c->blur + (size_t)y * c->blurStride + x * 4;
        uint8_t *d = c->dst + (size_t)y * c->dstStride + x * 4;
        for (int ch = 0; ch < 3; ch++) {
            double orig = (double)s[ch];
            double blur = (double)bl[ch];
            double val = orig + localAmount * (orig - blur);
            d[ch] = fo_clampf(val);
        }
        d[3] = s[3];
    }
}

/* effects.go:49-90 */
int fo_adaptive_sharpen(const uint8_t *src, int srcStride, int w, int h, double strength,
                        uint8_t *dst, int dstStride) {
    if (strength <= 0) return 1;
    if (strength > 1) strength = 1;
    if (w < 3 || h < 3) return 1;
    uint8_t *blurred = (uint8_t *)malloc((size_t)w * h * 4);
    fo_blur3x3(src, srcStride, w, h, blurred, w * 4);
    for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * dstStride, src + (size_t)y * srcStride, (size_t)w * 4);
    fo_fx_ctx c;
    c.src = src; c.srcStride = srcStride; c.blur = blurred; c.blurStride = w * 4;
    c.dst = dst; c.dstStride = dstStride; c.w = w; c.h = h;
    c.amount = 1.0 + strength * 2.0;
    parallel_do(1, h - 1, adaptive_row, &c);
    free(blurred);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * resize.go
 * ---------------------------------------------------------------------------------------- */

/* resize.go:55-69 */
double fo_lanczos_kernel(double x) {
    const double lanczosA = 3.0;
    if (x == 0) return 1.0;
    if (x < 0) x = -x;
    if (x >= lanczosA) return 0.0;
    double xpi = x * M_PI;
    return (lanczosA * sin(xpi) * sin(xpi / lanczosA)) / (xpi * xpi);
}

/* resize.go:81-85 / 125-129 */
static void lanczos_ratio_support(int srcSize, int dstSize, double *ratio, double *support) {
    *ratio = (double)srcSize / (double)dstSize;
    *support = 3.0;
    if (*ratio > 1) *support = 3.0 * *ratio;
}

static void lanczos_span(int d, int srcSize, double ratio, double support, double *center, int *left, int *right) {
    *center = ((double)d + 0.5) * ratio - 0.5; /* resize.go:169-178 */
    *left = (int)ceil(*center - support);
    *right = (int)floor(*center + support);
    if (*left < 0) *left = 0;
    if (*right >= srcSize) *right = srcSize - 1;
}

int fo_lanczos_weights_cap(int dstSize, int srcSize) {
    double ratio, support;
    lanczos_ratio_support(srcSize, dstSize, &ratio, &support);
    long long cap = 0;
    for (int d = 0; d < dstSize; d++) {
        double center;
        int left, right;
        lanczos_span(d, srcSize, ratio, support, &center, &left, &right);
        if (right >= left) cap += right - left + 1;
    }
    return (int)cap;
}

/* resize.go:164-197 */
int fo_lanczos_weights(int dstSize, int srcSize, int *start, int *index, double *weight) {
    double ratio, support;
    lanczos_ratio_support(srcSize, dstSize, &ratio, &support);
    double filterScale = fmax(ratio, 1.0);
    int n = 0;
    for (int d = 0; d < dstSize; d++) {
        double center;
        int left, right;
        lanczos_span(d, srcSize, ratio, support, &center, &left, &right);
        start[d] = n;
        double wsum = 0.0;
        int first = n;
        for (int s = left; s <= right; s++) {
            double w = fo_lanczos_kernel(((double)s - center) / filterScale);
            if (w != 0) {
                wsum += w;
                index[n] = s;
                weight[n] = w;
                n++;
            }
        }
        if (wsum != 0)
            for (int i = first; i < n; i++) weight[i] /= wsum;
    }
    start[dstSize] = n;
    return n;
}

typedef struct {
    const uint8_t *src;
    int srcStride;
    uint8_t *dst;
    int dstStride;
    int outer; /* dstW for H (loop over dx inside a row); dstH for V (loop over dy inside a column) */
    const int *start, *index;
    const double *weight;
} fo_rs_ctx;

/* resize.go:89-115 — one row */
static void resize_h_row(int y, void *vctx) {
    fo_rs_ctx *c = (fo_rs_ctx *)vctx;
    for (int dx = 0; dx < c->outer; dx++) {
        double r = 0, g = 0, b = 0, a = 0;
        for (int t = c->start[dx]; t < c->start[dx + 1]; t++) {
            const uint8_t *p = c->src + (size_t)y * c->srcStride + c->index[t] * 4;
            double sa = (double)p[3];
            double w = c->weight[t];
            double aw = sa * w;
            r += (double)p[0] * aw;
            g += (double)p[1] * aw;
            b += (double)p[2] * aw;
            a += aw;
        }
        if (a > 0.5) {
            uint8_t *q = c->dst + (size_t)y * c->dstStride + dx * 4;
            double inv = 1.0 / a;
            q[0] = fo_clampf(r * inv);
            q[1] = fo_clampf(g * inv);
            q[2] = fo_clampf(b * inv);
            q[3] = fo_clampf(a);
        }
    }
}

/* resize.go:133-158 — one column */
static void resize_v_col(int x, void *vctx) {
    fo_rs_ctx *c = (fo_rs_ctx *)vctx;
    for (int dy = 0; dy < c->outer; dy++) {
        double r = 0, g = 0, b = 0, a = 0;
        for (int t = c->start[dy]; t < c->start[dy + 1]; t++) {
            const uint8_t *p = c->src + (size_t)c->index[t] * c->srcStride + x * 4;
            double sa = (double)p[3];
            double w = c->weight[t];
            double aw = sa * w;
            r += (double)p[0] * aw;
            g += (double)p[1] * aw;
            b += (double)p[2] * aw;
            a += aw;
        }
        if (a > 0.5) {
            uint8_t *q = c->dst + (size_t)dy * c->dstStride + x * 4;
            double inv = 1.0 / a;
            q[0] = fo_clampf(r * inv);
            q[1] = fo_clampf(g * inv);
            q[2] = fo_clampf(b * inv);
            q[3] = fo_clampf(a);
        }
    }
}

/* resize.go:77-118 */
void fo_resize_h(const uint8_t *src, int srcStride, int srcW, int srcH,
                 uint8_t *dst, int dstStride, int dstW) {
    int cap = fo_lanczos_weights_cap(dstW, srcW);
    int *start = (int *)malloc(sizeof(int) * (size_t)(dstW + 1));
    int *index = (int *)malloc(sizeof(int) * (size_t)(cap + 1));
    double *weight = (double *)malloc(sizeof(double) * (size_t)(cap + 1));
    fo_lanczos_weights(dstW, srcW, start, index, weight);
    fo_rs_ctx c;
    c.src = src; c.srcStride = srcStride; c.dst = dst; c.dstStride = dstStride; c.outer = dstW;
    c.start = start; c.index = index; c.weight = weight;
    parallel_do(0, srcH, resize_h_row, &c);
    free(start); free(index); free(weight);
}

/* resize.go:121-161 */
void fo_resize_v(const uint8_t *src, int srcStride, int srcW, int srcH,
                 uint8_t *dst, int dstStride, int dstH) {
    int cap = fo_lanczos_weights_cap(dstH, srcH);
    int *start = (int *)malloc(sizeof(int) * (size_t)(dstH + 1));
    int *index = (int *)malloc(sizeof(int) * (size_t)(cap + 1));
    double *weight = (double *)malloc(sizeof(double) * (size_t)(cap + 1));
    fo_lanczos_weights(dstH, srcH, start, index, weight);
    fo_rs_ctx c;
    c.src = src; c.srcStride = srcStride; c.dst = dst; c.dstStride = dstStride; c.outer = dstH;
    c.start = start; c.index = index; c.weight = weight;
    parallel_do(0, srcW, resize_v_col, &c);
    free(start); free(index); free(weight);
}

/* resize.go:37-53 */
int fo_lanczos_resize(const uint8_t *src, int srcStride, int srcW, int srcH,
                      uint8_t *dst, int dstStride, int dstW, int dstH) {
    if (srcW <= 0 || srcH <= 0 || dstW <= 0 || dstH <= 0) return 1;
    if (srcW == dstW && srcH == dstH) {
        for (int y = 0; y < srcH; y++) memcpy(dst + (size_t)y * dstStride, src + (size_t)y * srcStride, (size_t)srcW * 4);
        return 0;
    }
    uint8_t *tmp = (uint8_t *)calloc((size_t)dstW * srcH, 4);
    fo_resize_h(src, srcStride, srcW, srcH, tmp, dstW * 4, dstW);
    fo_resize_v(tmp, dstW * 4, dstW, srcH, dst, dstStride, dstH);
    free(tmp);
    return 0;
}

/* resize.go:12-32 */
int fo_smart_resize_dims(int srcW, int srcH, int maxW, int maxH, int *dstW, int *dstH) {
    if (maxW <= 0) maxW = srcW;
    if (maxH <= 0) maxH = srcH;
    if (srcW <= maxW && srcH <= maxH) {
        *dstW = srcW;
        *dstH = srcH;
        return 1;
    }
    double ratio = fmin((double)maxW / (double)srcW, (double)maxH / (double)srcH);
    *dstW = (int)fmax(1, round((double)srcW * ratio));
    *dstH = (int)fmax(1, round((double)srcH * ratio));
    return 0;
}

/* ---- §8(f1): convertToNRGBA on decoded images (convert.go:34-64) ---------------------------------
 * The reference walks img.At(x, y).RGBA() for every pixel.  For the two concrete types jpeg.Decode returns this
 * is Go standard-library arithmetic (image/ycbcr.go, image/color/ycbcr.go — Go 1.25.5 per go.mod:3, NOT under
 * /root/reference; restated here from the published source):
 *   (*image.YCbCr).At     -> YCbCrAt: Y[YOffset(x,y)], Cb/Cr[COffset(x,y)], COffset per subsample ratio
 *   color.YCbCr.RGBA()    -> yy1 = Y*0x10101; cb1 = Cb-128; cr1 = Cr-128;
 *                            r = yy1 + 91881*cr1; g = yy1 - 22554*cb1 - 46802*cr1; b = yy1 + 116130*cb1;
 *                            each: if uint32(v)&0xff000000 == 0 { v >>= 8 } else { v = ^(v>>31) & 0xffff }; a = 0xffff
 *   convertToNRGBA, a == 0xffff branch (convert.go:48-53): uint8(v >> 8)
 * which composes to the 8-bit color.YCbCrToRGB result.  Rect.Min is (0,0) for decoded images. */
static int fo_c_offset(int ratio, int x, int y, int cStride) {
    switch (ratio) {            /* image.YCbCrSubsampleRatio constants, in Go's order */
        case 1: return y * cStride + x / 2;        /* 4:2:2 */
        case 2: return (y / 2) * cStride + x / 2;  /* 4:2:0 */
        case 3: return (y / 2) * cStride + x;      /* 4:4:0 */
        case 4: return y * cStride + x / 4;        /* 4:1:1 */
        case 5: return (y / 2) * cStride + x / 4;  /* 4:1:0 */
        default: return y * cStride + x;           /* 4:4:4 */
    }
}

static uint8_t fo_ycc_channel(int32_t v) {
    uint32_t c16;
    if (((uint32_t)v & 0xff000000u) == 0) c16 = (uint32_t)(v >> 8);
    else c16 = (uint32_t)(~(v >> 31)) & 0xffffu;   /* arithmetic shift: -1 for negatives, 0 for overflow */
    return (uint8_t)(c16 >> 8);
}

int fo_ycbcr_to_nrgba(const uint8_t *yp, int yStride, const uint8_t *cb, const uint8_t *cr, int cStride,
                      int w, int h, int ratio, uint8_t *dst, int dstStride) {
    if (ratio < 0 || ratio > 5) return -1;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            int ci = fo_c_offset(ratio, x, y, cStride);
            int32_t yy1 = (int32_t)yp[y * yStride + x] * 0x10101;
            int32_t cb1 = (int32_t)cb[ci] - 128;
            int32_t cr1 = (int32_t)cr[ci] - 128;
            uint8_t *o = dst + (size_t)y * dstStride + (size_t)x * 4;
            o[0] = fo_ycc_channel(yy1 + 91881 * cr1);
            o[1] = fo_ycc_channel(yy1 - 22554 * cb1 - 46802 * cr1);
            o[2] = fo_ycc_channel(yy1 + 116130 * cb1);
            o[3] = 0xff;
        }
    }
    return 0;
}

/* (*image.Gray).At -> color.Gray{Y}.RGBA() = (y|y<<8) x3, 0xffff; convert.go:48-53 takes the high byte. */
void fo_gray_to_nrgba(const uint8_t *g, int gStride, int w, int h, uint8_t *dst, int dstStride) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint8_t v = g[y * gStride + x];
            uint8_t *o = dst + (size_t)y * dstStride + (size_t)x * 4;
            o[0] = o[1] = o[2] = v;
            o[3] = 0xff;
        }
}

/* ---- convertToNRGBA (convert.go:34-64) for the other concrete types image/png and image/jpeg decode into -----
 * Step 1 is Go's At(x, y).RGBA() of each type (image/image.go, image/color/color.go, image/color/ycbcr.go — Go
 * 1.25.5 standard library, not vendored; restated from the published source), step 2 is convert.go:42-60 verbatim:
 * a == 0 -> zeros; a == 0xffff -> channel >> 8; else uint8(((c * 0xffff) / a) >> 8), alpha a >> 8, in uint32
 * arithmetic with Go's truncating uint8() conversion.  Rect.Min == (0, 0) as every decoder returns.
 *   1 *image.RGBA     Pix R,G,B,A 8-bit premultiplied: color.RGBA.RGBA() = v | v<<8 per field
 *   2 *image.RGBA64   Pix big-endian 16-bit R,G,B,A premultiplied: the fields themselves
 *   3 *image.NRGBA64  big-endian 16-bit, straight alpha: c * a / 0xffff per colour field, a
 *   4 *image.Gray16   big-endian 16-bit Y: (y, y, y, 0xffff)
 *   5 *image.CMYK     Pix C,M,Y,K: w = 0xffff - k*0x101; (0xffff - c*0x101) * w / 0xffff; alpha 0xffff
 *   6 *image.Paletted Pix = index; pal16 holds Palette[i].RGBA() as 4 x uint16 per entry (the caller evaluates the
 *                     color.Color interface; values are <= 0xffff by contract).  Go panics on an index >= len(Palette):
 *                     the restatement returns -2 instead. */
static void fo_convert_px(uint32_t r, uint32_t g, uint32_t b, uint32_t a, uint8_t *o) {
    if (a == 0) {
        o[0] = o[1] = o[2] = o[3] = 0;
    } else if (a == 0xffff) {
        o[0] = (uint8_t)(r >> 8); o[1] = (uint8_t)(g >> 8); o[2] = (uint8_t)(b >> 8); o[3] = 0xff;
    } else {
        o[0] = (uint8_t)(((r * 0xffffu) / a) >> 8);
        o[1] = (uint8_t)(((g * 0xffffu) / a) >> 8);
        o[2] = (uint8_t)(((b * 0xffffu) / a) >> 8);
        o[3] = (uint8_t)(a >> 8);
    }
}

static uint32_t fo_be16(const uint8_t *p) { return ((uint32_t)p[0] << 8) | p[1]; }

int fo_convert_to_nrgba(int fmt, const uint8_t *pix, int stride, int w, int h, const uint16_t *pal16, int ncolors,
                        uint8_t *dst, int dstStride) {
    if (fmt < 1 || fmt > 6) return -1;
    if (fmt == 6 && (!pal16 || ncolors < 1 || ncolors > 256)) return -1;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            uint32_t r, g, b, a;
            const uint8_t *row = pix + (size_t)y * stride;
            switch (fmt) {
                case 1: {
                    const uint8_t *p = row + (size_t)x * 4;
                    r = p[0] | ((uint32_t)p[0] << 8); g = p[1] | ((uint32_t)p[1] << 8);
                    b = p[2] | ((uint32_t)p[2] << 8); a = p[3] | ((uint32_t)p[3] << 8);
                } break;
                case 2: {
                    const uint8_t *p = row + (size_t)x * 8;
                    r = fo_be16(p); g = fo_be16(p + 2); b = fo_be16(p + 4); a = fo_be16(p + 6);
                } break;
                case 3: {
                    const uint8_t *p = row + (size_t)x * 8;
                    a = fo_be16(p + 6);
                    r = fo_be16(p) * a / 0xffffu; g = fo_be16(p + 2) * a / 0xffffu; b = fo_be16(p + 4) * a / 0xffffu;
                } break;
                case 4: {
                    r = g = b = fo_be16(row + (size_t)x * 2);
                    a = 0xffff;
                } break;
                case 5: {
                    const uint8_t *p = row + (size_t)x * 4;
                    uint32_t wk = 0xffffu - (uint32_t)p[3] * 0x101u;
                    r = (0xffffu - (uint32_t)p[0] * 0x101u) * wk / 0xffffu;
                    g = (0xffffu - (uint32_t)p[1] * 0x101u) * wk / 0xffffu;
                    b = (0xffffu - (uint32_t)p[2] * 0x101u) * wk / 0xffffu;
                    a = 0xffff;
                } break;
                default: {
                    int i = row[x];
                    if (i >= ncolors) return -2;
                    r = pal16[4 * i]; g = pal16[4 * i + 1]; b = pal16[4 * i + 2]; a = pal16[4 * i + 3];
                } break;
            }
            fo_convert_px(r, g, b, a, dst + (size_t)y * dstStride + (size_t)x * 4);
        }
    }
    return 0;
}

/* ---- §8(f2): Analyze (analyze.go:26-176) — the measured part: one full scan + two sampled scans ------------
 * Sequential, source order, binary64 unfused — exactly as the Go loops.  math.Log2 is Go's own
 * (Frexp; frac == 0.5 -> exp-1; else Log(frac)*(1/Ln2) + exp), restated with libm's log: entropy may differ from Go
 * in the last bits (both logs are < 1 ulp), which is why the GPU parity bar on Entropy is a tolerance.
 * The recommendations (analyze.go:183-232) are host logic on these numbers and are restated too. */

static double fo_go_log2(double x) {
    int e;
    double frac = frexp(x, &e);
    if (frac == 0.5) return (double)(e - 1);
    return log(frac) * (1.0 / 0.693147180559945309417232121458176568) + (double)e;
}

static double fo_lum_at(const uint8_t *pix, int stride, int x, int y) {   /* analyze.go:178-181 */
    const uint8_t *p = pix + (size_t)y * stride + (size_t)x * 4;
    return 0.299 * (double)p[0] + 0.587 * (double)p[1] + 0.114 * (double)p[2];
}

static int fo_cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* Numeric values are Go's (types.go:36-42, 59-70): Format Auto=0 JPEG=1 PNG=2; Quality Balanced=0 Lossless=1 Ultra=2 High=3 Aggressive=4 */
#define FO_JPEG 1
#define FO_PNG 2
#define FO_BALANCED 0
#define FO_HIGH 3
#define FO_AGGRESSIVE 4
void fo_recommend(fo_image_stats *st) {
    /* recommendFormat analyze.go:183-194 */
    if (st->has_alpha) st->recommended_format = FO_PNG;
    else if (st->unique_colors <= 256) st->recommended_format = FO_PNG;
    else if (st->edge_density > 0.3 && st->unique_colors < 1000) st->recommended_format = FO_PNG;
    else st->recommended_format = FO_JPEG;
    /* recommendQuality analyze.go:196-207 */
    if (st->entropy > 6 && st->edge_density < 0.15) st->recommended_quality = FO_BALANCED;
    else if (st->entropy < 4) st->recommended_quality = FO_AGGRESSIVE;
    else if (st->edge_density > 0.25) st->recommended_quality = FO_HIGH;
    else st->recommended_quality = FO_BALANCED;
    /* estimateCompression analyze.go:209-232 */
    if (st->recommended_format == FO_PNG) {
        if (st->unique_colors <= 256) st->estimated_compression = 5.0 + (256 - (double)st->unique_colors) / 50;
        else if (st->is_grayscale) st->estimated_compression = 3.0;
        else st->estimated_compression = 2.0;
    } else {
        double base = 10.0;
        if (st->entropy > 7) base = 5.0;
        else if (st->entropy > 5) base = 8.0;
        if (st->edge_density > 0.2) base *= 0.7;
        st->estimated_compression = base;
    }
}

void fo_analyze(const uint8_t *pix, int stride, int w, int h, fo_image_stats *st) {
    memset(st, 0, sizeof *st);
    st->width = w;
    st->height = h;
    if (w == 0 || h == 0) return;                       /* analyze.go:36-38 */
    double bright = 0.0;
    const int max_sample = 50000;
    long long step = 1;
    if ((long long)w * h > max_sample) step = (long long)w * h / max_sample;
    /* the map[uint32]struct{} capped at 1024 entries: collect the sampled keys in scan order, stop at 1024 distinct */
    uint32_t *keys = (uint32_t *)malloc(sizeof(uint32_t) * 1024);
    int nkeys = 0;
    int all_gray = 1, has_alpha = 0;
    long long idx = 0;
    for (int y = 0; y < h; y++) {
        const uint8_t *row = pix + (size_t)y * stride;
        for (int x = 0; x < w; x++) {
            uint8_t r = row[4 * x], g = row[4 * x + 1], b = row[4 * x + 2], a = row[4 * x + 3];
            double lum = 0.299 * (double)r + 0.587 * (double)g + 0.114 * (double)b;
            bright += lum;
            st->histogram[(int)(lum + 0.5)] += 1.0;
            if (a < 255) has_alpha = 1;
            if (r != g || g != b) all_gray = 0;
            if (idx % step == 0 && nkeys < 1024) {
                uint32_t key = (uint32_t)r << 24 | (uint32_t)g << 16 | (uint32_t)b << 8 | (uint32_t)a;
                int found = 0;
                for (int k = 0; k < nkeys; k++) if (keys[k] == key) { found = 1; break; }
                if (!found) keys[nkeys++] = key;
            }
            idx++;
        }
    }
    (void)fo_cmp_u32;
    free(keys);
    double n = (double)((long long)w * h);
    st->has_alpha = has_alpha;
    st->is_grayscale = all_gray;
    st->unique_colors = nkeys;
    st->mean_brightness = bright / n;
    int step_y = (int)fmax(1, ceil((double)h / 100));
    int step_x = (int)fmax(1, ceil((double)w / 100));
    double var_sum = 0.0, mean = st->mean_brightness;
    int samples = 0;
    for (int y = 0; y < h; y += step_y)
        for (int x = 0; x < w; x += step_x) {
            double d = fo_lum_at(pix, stride, x, y) - mean;
            var_sum += d * d;
            samples++;
        }
    if (samples > 0) st->contrast = sqrt(var_sum / (double)samples);
    double entropy = 0.0;                                /* computeEntropy analyze.go:116-128 */
    for (int i = 0; i < 256; i++)
        if (st->histogram[i] > 0) {
            double p = st->histogram[i] / n;
            entropy -= p * fo_go_log2(p);
        }
    st->entropy = entropy;
    if (w >= 3 && h >= 3) {                              /* computeEdgeDensity analyze.go:131-176 */
        int sx = (int)fmax(1, (double)w / 200), sy = (int)fmax(1, (double)h / 200);
        int edges = 0, total = 0;
        for (int y = 1; y < h - 1; y += sy)
            for (int x = 1; x < w - 1; x += sx) {
#define L(xx, yy) fo_lum_at(pix, stride, (xx), (yy))
                double gx = L(x + 1, y - 1) - L(x - 1, y - 1) + 2 * L(x + 1, y) - 2 * L(x - 1, y) + L(x + 1, y + 1) - L(x - 1, y + 1);
                double gy = L(x - 1, y + 1) - L(x - 1, y - 1) + 2 * L(x, y + 1) - 2 * L(x, y - 1) + L(x + 1, y + 1) - L(x + 1, y - 1);
#undef L
                if (sqrt(gx * gx + gy * gy) > 30.0) edges++;
                total++;
            }
        if (total > 0) st->edge_density = (double)edges / (double)total;
    }
    fo_recommend(st);
}

/* ---- §8(f4): ApplyOrientation (exif.go:176-203) through the loops of convert.go:186-256 -------------------------
 * dst must hold the oriented image (w x h, or h x w for orientations 5-8).  Returns 1 when the reference returns
 * its input (orientations 1, 0, unknown), dst untouched. */
static void fo_perm(const uint8_t *src, int ss, int w, int h, uint8_t *dst, int ds, int kind) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            size_t so = (size_t)y * ss + (size_t)x * 4, d;
            switch (kind) {
                case 0: d = (size_t)x * ds + (size_t)(h - 1 - y) * 4; break;           /* rotate90CW  convert.go:186-198 */
                case 1: d = (size_t)(h - 1 - y) * ds + (size_t)(w - 1 - x) * 4; break; /* rotate180    :201-213 */
                case 2: d = (size_t)(w - 1 - x) * ds + (size_t)y * 4; break;           /* rotate270CW  :216-226 */
                case 3: d = (size_t)y * ds + (size_t)(w - 1 - x) * 4; break;           /* flipH        :229-241 */
                default: d = (size_t)(h - 1 - y) * ds + (size_t)x * 4; break;          /* flipV        :244-255 */
            }
            memcpy(dst + d, src + so, 4);
        }
}

int fo_apply_orientation(const uint8_t *src, int ss, int w, int h, int orient, uint8_t *dst, int ds) {
    switch (orient) {
        case 2: fo_perm(src, ss, w, h, dst, ds, 3); return 0;
        case 3: fo_perm(src, ss, w, h, dst, ds, 1); return 0;
        case 4: fo_perm(src, ss, w, h, dst, ds, 4); return 0;
        case 6: fo_perm(src, ss, w, h, dst, ds, 0); return 0;
        case 8: fo_perm(src, ss, w, h, dst, ds, 2); return 0;
        case 5: case 7: {   /* rotate, then flipH of the rotated (h x w) image: exif.go:187-196 */
            uint8_t *tmp = (uint8_t *)malloc((size_t)h * 4 * (size_t)w + 4);
            fo_perm(src, ss, w, h, tmp, h * 4, orient == 5 ? 2 : 0);
            fo_perm(tmp, h * 4, h, w, dst, ds, 3);
            free(tmp);
            return 0;
        }
        default: return 1;
    }
}

/* ---- §8(f3): applyPalette + palettedToNRGBA (targetsize.go:479-545) ----------------------------------------------
 * palette: ncolors NRGBA entries (A = 255 as medianCut produces them, so c.RGBA()>>8 is the 8-bit channel).
 * The reference's map is a memo of the same search and is omitted.  idx / out may be NULL. */
void fo_apply_palette(const uint8_t *src, int ss, int w, int h, const uint8_t *palette, int ncolors,
                      uint8_t *idx, int idxStride, uint8_t *out, int outStride) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t *p = src + (size_t)y * ss + (size_t)x * 4;
            int bestIdx = 0, bestDist = 2147483647;                    /* math.MaxInt32 */
            for (int i = 0; i < ncolors; i++) {
                int dr = (int)p[0] - (int)palette[4 * i], dg = (int)p[1] - (int)palette[4 * i + 1], db = (int)p[2] - (int)palette[4 * i + 2];
                int dist = dr * dr + dg * dg + db * db;
                if (dist < bestDist) { bestDist = dist; bestIdx = i; }
            }
            if (idx) idx[(size_t)y * idxStride + x] = (uint8_t)bestIdx;
            if (out) memcpy(out + (size_t)y * outStride + (size_t)x * 4, palette + 4 * bestIdx, 4);   /* targetsize.go:529-538, A = 255 */
        }
}
