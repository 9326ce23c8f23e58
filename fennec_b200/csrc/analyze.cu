// analyze.cu — SURVEY §8(f2): the scans behind Analyze (analyze.go:26-176) on a device-resident NRGBA image.
//
// The reference makes one full pass (luminance histogram, brightness sum, alpha / grayscale tests, a sampled
// colour set capped at 1024 entries) and two sampled passes (contrast on a <= 100x100 grid, Sobel edge density on a
// <= ~200x200 grid).  Here (K9a-c are block roles of ONE launch, analyze_all_kernel):
//   K9a analyze_scan_role      the full pass: 4 B/px read once, HBM-bound.  Luminance x1000 is the exact integer
//                              L = 299R + 587G + 114B (two IDP); the histogram bin int(lum + 0.5) is (L + 500) / 1000
//                              except when (L + 500) % 1000 == 0, where the mathematically exact value sits ON the
//                              bin edge and the reference's FP64 rounding decides — only then the FP64 expression
//                              of analyze.go:63 is evaluated (1 pixel in 1000).  Per-warp shared-memory
//                              histograms, integer sums: the result does not depend on the summation order.
//   K9b analyze_sample_role    the colour set: sample k is pixel k*step in scan order (analyze.go:45-48, 73-76);
//                              keys go into an open-addressing table with 64-bit CAS, distinct insertions are counted
//                              (the map stops growing at 1024, so UniqueColors = min(distinct, 1024)).
//   K9c analyze_edge_role      Sobel on the sample grid in the reference's FP64 expression order (analyze.go:148-156);
//                              integer counts.
//   K9d analyze_contrast_kernel  sum((lum - mean)^2) on the grid, one block, fixed reduction tree (deterministic).
// MeanBrightness: the reference adds 8 M doubles sequentially; sum(L)/1000/n is the same quantity without the
// accumulated rounding (agreement ~1e-12 relative; the bar in tests/ is 1e-9).  Entropy, sqrt and the ratios are
// finished on the host from the raw record (api.cu: fb_analyze_finish), like every other table/rule of the
// host side of the boundary.
#include "common.cuh"

namespace fb {

namespace {

constexpr int kScanThreads = 256;

__device__ __forceinline__ uint32_t luma1000(uint32_t px) {
    return __dp2a_lo(299u | (587u << 16), px, __dp2a_hi(114u, px, 0u));
}

__device__ __forceinline__ double lum_fp64(uint32_t px) {   // analyze.go:63 / 178-181, unfused, left to right
    const double r = (double)(px & 0xFF), g = (double)((px >> 8) & 0xFF), b = (double)((px >> 16) & 0xFF);
    return __dadd_rn(__dadd_rn(__dmul_rn(0.299, r), __dmul_rn(0.587, g)), __dmul_rn(0.114, b));
}

__device__ __forceinline__ int lum_bin(uint32_t px, uint32_t L) {
    const uint32_t q = (L + 500u) / 1000u;
    if ((L + 500u) - q * 1000u != 0u) return (int)q;
    return (int)__dadd_rn(lum_fp64(px), 0.5);   // on the edge: the reference's own rounding decides
}

__device__ __forceinline__ void analyze_scan_role(unsigned int (*hist)[256], int bx, const uint8_t *imgs, long long imgStride,
                                                  int rowStride, int w, int h, int rowsPerBlock, AnalyzeRaw *raw, int vecOK) {
    const int img = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (kScanThreads / 32) * 256; i += kScanThreads) (&hist[0][0])[i] = 0u;
    __syncthreads();
    const uint8_t *base = imgs + (long long)img * imgStride;
    const int y0 = bx * rowsPerBlock, y1 = min(y0 + rowsPerBlock, h);
    unsigned long long sumL = 0;
    uint32_t sum32 = 0;                       // per-sweep partial of sumL: <= 16 pixels x 255 000 between flushes
    uint32_t andAll = 0xFFFFFFFFu, orColour = 0;
    unsigned int *myHist = hist[warp];
    auto px1 = [&](uint32_t v) {
        const uint32_t L = luma1000(v);
        sum32 += L;
        atomicAdd(&myHist[lum_bin(v, L)], 1u);
        andAll &= v;                          // alpha < 255 somewhere  <=>  the top byte of the AND is not 0xFF
        orColour |= v ^ (v >> 8);             // bytes 0, 1 = r^g, g^b (masked once at the end)
    };
    for (int y = y0; y < y1; y++) {
        const uint8_t *row = base + (long long)y * rowStride;
        if (vecOK) {
            // four independent 128-bit loads in flight per thread (one 4096-pixel sweep of the row per iteration)
            for (int xb = threadIdx.x * 4; xb < w; xb += kScanThreads * 16) {
                uint4 q[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int x = xb + u * kScanThreads * 4;
                    q[u] = (x + 4 <= w) ? ld_nc_u128(row + (long long)x * 4) : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int x = xb + u * kScanThreads * 4;
                    if (x + 4 <= w) {
                        px1(q[u].x); px1(q[u].y); px1(q[u].z); px1(q[u].w);
                    } else {
                        for (int i = 0; x + i < w; i++) px1(ld_nc_u32(row + (long long)(x + i) * 4));
                    }
                }
                sumL += sum32;
                sum32 = 0;
            }
        } else {
            for (int x = threadIdx.x; x < w; x += kScanThreads) {
                px1(ld_nc_u32(row + (long long)x * 4));
                sumL += sum32;
                sum32 = 0;
            }
        }
    }
    uint32_t orAlphaLow = ~andAll & 0xFF000000u;
    orColour &= 0x0000FFFFu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sumL += __shfl_xor_sync(0xffffffffu, sumL, o);
        orAlphaLow |= __shfl_xor_sync(0xffffffffu, orAlphaLow, o);
        orColour |= __shfl_xor_sync(0xffffffffu, orColour, o);
    }
    AnalyzeRaw *r = raw + img;
    if (lane == 0) {
        atomicAdd(&r->sumL, sumL);
        if (orAlphaLow) r->hasAlpha = 1u;   // benign race: every writer stores 1
        if (orColour) r->hasColour = 1u;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 256; b += kScanThreads) {
        unsigned int c = 0;
#pragma unroll
        for (int k = 0; k < kScanThreads / 32; k++) c += hist[k][b];
        if (c) atomicAdd(&r->hist[b], c);
    }
}

__device__ __forceinline__ void analyze_sample_role(int bx, const uint8_t *imgs, long long imgStride, int rowStride, int w,
                                                    long long step, int nSamples, unsigned long long *tables, int tableMask,
                                                    AnalyzeRaw *raw) {
    const int k = bx * blockDim.x + threadIdx.x, img = blockIdx.y;
    if (k >= nSamples) return;
    const long long idx = (long long)k * step;
    const int y = (int)(idx / w), x = (int)(idx - (long long)y * w);
    const uint32_t v = ld_nc_u32(imgs + (long long)img * imgStride + (long long)y * rowStride + (long long)x * 4);
    // analyze.go:74: r<<24 | g<<16 | b<<8 | a (any injective key gives the same count; kept for readability)
    const uint32_t key = __byte_perm(v, 0, 0x0123);
    const unsigned long long tagged = (unsigned long long)key | (1ull << 32);   // 0 = empty slot
    unsigned long long *table = tables + (size_t)img * (tableMask + 1);
    uint32_t slot = (key * 2654435761u) & (uint32_t)tableMask;
    for (;;) {
        const unsigned long long prev = atomicCAS(&table[slot], 0ull, tagged);
        if (prev == 0ull) { atomicAdd(&raw[img].uniqueSampled, 1u); return; }
        if (prev == tagged) return;
        slot = (slot + 1) & (uint32_t)tableMask;
    }
}

__device__ __forceinline__ void analyze_edge_role(int bx, const uint8_t *imgs, long long imgStride, int rowStride, int w, int h,
                                                  int sx, int sy, int nx, int ny, AnalyzeRaw *raw) {
    const int i = bx * blockDim.x + threadIdx.x, img = blockIdx.y;
    bool edge = false;
    if (i < nx * ny) {
        const int x = 1 + (i % nx) * sx, y = 1 + (i / nx) * sy;
        const uint8_t *b = imgs + (long long)img * imgStride;
        auto L = [&](int xx, int yy) { return lum_fp64(ld_nc_u32(b + (long long)yy * rowStride + (long long)xx * 4)); };
        // analyze.go:148-154, left to right
        double gx = __dadd_rn(L(x + 1, y - 1), -L(x - 1, y - 1));
        gx = __dadd_rn(gx, __dmul_rn(2.0, L(x + 1, y)));
        gx = __dadd_rn(gx, -__dmul_rn(2.0, L(x - 1, y)));
        gx = __dadd_rn(gx, L(x + 1, y + 1));
        gx = __dadd_rn(gx, -L(x - 1, y + 1));
        double gy = __dadd_rn(L(x - 1, y + 1), -L(x - 1, y - 1));
        gy = __dadd_rn(gy, __dmul_rn(2.0, L(x, y + 1)));
        gy = __dadd_rn(gy, -__dmul_rn(2.0, L(x, y - 1)));
        gy = __dadd_rn(gy, L(x + 1, y + 1));
        gy = __dadd_rn(gy, -L(x + 1, y - 1));
        edge = __dsqrt_rn(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy))) > 30.0;
    }
    const unsigned int m = __ballot_sync(0xffffffffu, edge);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&raw[img].edges, (unsigned int)__popc(m));
}

// One launch for the three independent passes: blocks [0, scanBlocks) run the full scan, the next sampleBlocks the
// colour sampling, the rest the Sobel grid — so the two latency-bound sampled passes run underneath the
// bandwidth-bound scan instead of after it (they were 25 of the 59 us per 4 images as separate launches).
struct AnalyzeArgs {
    const uint8_t *imgs;
    long long imgStride;
    int rowStride, w, h, vecOK;
    int rowsPerBlock, scanBlocks, sampleBlocks;
    long long sampleStep;
    int nSamples, tableMask;
    unsigned long long *tables;
    int edgeX, edgeY, edgeNx, edgeNy;
    AnalyzeRaw *raw;
};

__global__ void __launch_bounds__(kScanThreads) analyze_all_kernel(const AnalyzeArgs a) {
    __shared__ unsigned int hist[kScanThreads / 32][256];
    const int bx = blockIdx.x;
    if (bx < a.scanBlocks) analyze_scan_role(hist, bx, a.imgs, a.imgStride, a.rowStride, a.w, a.h, a.rowsPerBlock, a.raw, a.vecOK);
    else if (bx < a.scanBlocks + a.sampleBlocks)
        analyze_sample_role(bx - a.scanBlocks, a.imgs, a.imgStride, a.rowStride, a.w, a.sampleStep, a.nSamples, a.tables, a.tableMask, a.raw);
    else
        analyze_edge_role(bx - a.scanBlocks - a.sampleBlocks, a.imgs, a.imgStride, a.rowStride, a.w, a.h, a.edgeX, a.edgeY, a.edgeNx, a.edgeNy, a.raw);
}

constexpr int kContrastThreads = 1024;   // <= 10 000 grid samples per image: ~10 per thread, so the load latency is paid ~5 times, not 40
__global__ void __launch_bounds__(kContrastThreads) analyze_contrast_kernel(const uint8_t *imgs, long long imgStride, int rowStride, int w, int h,
                                                                            int sx, int sy, int nx, int ny, AnalyzeRaw *raw) {
    __shared__ double part[kContrastThreads];
    const int img = blockIdx.x;
    AnalyzeRaw *r = raw + img;
    // MeanBrightness (analyze.go:86) from the exact integer sum
    const double mean = __ddiv_rn(__ddiv_rn((double)r->sumL, 1000.0), (double)((long long)w * h));
    const uint8_t *b = imgs + (long long)img * imgStride;
    double acc = 0.0;
    const int n = nx * ny;
    for (int i0 = threadIdx.x; i0 < n; i0 += 2 * kContrastThreads) {   // two independent loads in flight
        const int i1 = i0 + kContrastThreads;
        const uint32_t v0 = ld_nc_u32(b + (long long)((i0 / nx) * sy) * rowStride + (long long)((i0 % nx) * sx) * 4);
        const uint32_t v1 = i1 < n ? ld_nc_u32(b + (long long)((i1 / nx) * sy) * rowStride + (long long)((i1 % nx) * sx) * 4) : 0u;
        const double d0 = __dadd_rn(lum_fp64(v0), -mean);
        acc = __dadd_rn(acc, __dmul_rn(d0, d0));
        if (i1 < n) {
            const double d1 = __dadd_rn(lum_fp64(v1), -mean);
            acc = __dadd_rn(acc, __dmul_rn(d1, d1));
        }
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = kContrastThreads / 2; o > 0; o >>= 1) {   // fixed tree: the result does not depend on scheduling
        if ((int)threadIdx.x < o) part[threadIdx.x] = __dadd_rn(part[threadIdx.x], part[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) r->varSum = part[0];
}

}  // namespace

size_t analyze_scratch_bytes(int w, int h, int n) {
    (void)w; (void)h;
    return (size_t)n * kAnalyzeTableSlots * sizeof(unsigned long long) + 1024;
}

// Grid steps of analyze.go:90-91 (contrast) and :139-140 (edges), and the sampling step of :45-48.
void analyze_steps(int w, int h, AnalyzeSteps *s) {
    s->contrastY = (int)fmax(1.0, ceil((double)h / 100));
    s->contrastX = (int)fmax(1.0, ceil((double)w / 100));
    s->contrastNy = (h + s->contrastY - 1) / s->contrastY;
    s->contrastNx = (w + s->contrastX - 1) / s->contrastX;
    s->edgeX = (int)fmax(1.0, (double)w / 200);
    s->edgeY = (int)fmax(1.0, (double)h / 200);
    s->edgeNx = (w >= 3 && h >= 3) ? (w - 2 + s->edgeX - 1) / s->edgeX : 0;   // x = 1, 1+sx, ... < w-1
    s->edgeNy = (w >= 3 && h >= 3) ? (h - 2 + s->edgeY - 1) / s->edgeY : 0;
    const long long px = (long long)w * h;
    s->sampleStep = px > 50000 ? px / 50000 : 1;
    s->nSamples = (int)((px + s->sampleStep - 1) / s->sampleStep);
}

int launch_analyze(cudaStream_t s, const uint8_t *imgs, long long imgStride, int rowStride, int w, int h, int n,
                   AnalyzeRaw *raw, void *scratch) {
    if (n <= 0 || w <= 0 || h <= 0) return FB_OK;
    AnalyzeSteps st;
    analyze_steps(w, h, &st);
    if (st.nSamples > kAnalyzeTableSlots / 2) { return FB_E_INVALID; }   // cannot happen: nSamples < 100001
    unsigned long long *tables = (unsigned long long *)scratch;
    FB_CUDA(cudaMemsetAsync(raw, 0, sizeof(AnalyzeRaw) * (size_t)n, s));
    FB_CUDA(cudaMemsetAsync(tables, 0, sizeof(unsigned long long) * (size_t)kAnalyzeTableSlots * n, s));
    const int vecOK = (((uintptr_t)imgs | (uintptr_t)imgStride | (uintptr_t)rowStride) & 15) == 0;
    // ~4 resident CTAs per SM across the batch, at least 4 rows each
    int blocksPerImg = (148 * 8 + n - 1) / n;
    int rowsPerBlock = (h + blocksPerImg - 1) / blocksPerImg;
    if (rowsPerBlock < 4) rowsPerBlock = 4;
    blocksPerImg = (h + rowsPerBlock - 1) / rowsPerBlock;
    AnalyzeArgs a;
    a.imgs = imgs; a.imgStride = imgStride; a.rowStride = rowStride; a.w = w; a.h = h; a.vecOK = vecOK;
    a.rowsPerBlock = rowsPerBlock; a.scanBlocks = blocksPerImg;
    a.sampleStep = st.sampleStep; a.nSamples = st.nSamples; a.sampleBlocks = (st.nSamples + kScanThreads - 1) / kScanThreads;
    a.tables = tables; a.tableMask = kAnalyzeTableSlots - 1;
    a.edgeX = st.edgeX; a.edgeY = st.edgeY; a.edgeNx = st.edgeNx; a.edgeNy = st.edgeNy;
    a.raw = raw;
    const int edgeBlocks = (st.edgeNx * st.edgeNy + kScanThreads - 1) / kScanThreads;
    analyze_all_kernel<<<dim3(a.scanBlocks + a.sampleBlocks + edgeBlocks, n), kScanThreads, 0, s>>>(a);
    int launches = 1;
    analyze_contrast_kernel<<<n, kContrastThreads, 0, s>>>(imgs, imgStride, rowStride, w, h, st.contrastX, st.contrastY, st.contrastNx, st.contrastNy, raw);
    launches++;
    FB_LAUNCHED(launches);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

}  // namespace fb
