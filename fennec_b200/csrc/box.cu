// box.cu — K2: variable-box mean downsample of all four channels, bit-exact with
// boxDownsample + averageBoxPixel (ssim.go:244-309).
//
// The reference sums uint8 values in float64 (exact integers), multiplies once by 1.0/count and
// rounds half away from zero.  Here the sums are integers, the single multiply is an IEEE binary64
// __dmul_rn and the rounding is clampf_dev, so every output byte is identical.  Box edges are
// int(float64(d) * ratio) with the reference's clamps (ssim.go:255-278); they are recomputed on the
// device with the same IEEE double multiply + truncation, so no tables cross the ABI.
//
// Mapping (HBM-bound: every source byte is read exactly once, fully coalesced): one CTA per
// (output row, chunk of output columns).  Phase 1 — each thread owns 4 adjacent source columns
// (one 128-bit load per row) and accumulates the vertical sums of the box rows as packed 2x16-bit
// integers; the column sums go to shared memory.  Phase 2 — one thread per output pixel adds the
// column sums of its box, multiplies by the reciprocal count in FP64 and writes 4 bytes.
#include "common.cuh"

#include <stdlib.h>
#include <vector>

namespace fb {

namespace {

constexpr int kSpanMax = 2048;   // source columns whose sums fit in shared memory (16 KB)
constexpr int kThreads = 256;

struct BoxParams {
    const uint8_t *src;
    uint8_t *dst;
    long long srcImgStride, dstImgStride;
    int srcRowStride, dstRowStride;
    int srcW, srcH, dstW, dstH;
    double xRatio, yRatio;
    int dxChunk;   // output columns per CTA
    int dyPerCta;  // output rows per CTA (small boxes: amortise the CTA over more source rows)
    int vecOK;
    // optional second batch in the same launch (SSIMFast downsamples both images of every pair): blockIdx.z >= nFirst
    const uint8_t *src2;
    uint8_t *dst2;
    long long src2ImgStride;
    int src2RowStride, nFirst;
};

// ssim.go:255-265 / 268-278
__device__ __forceinline__ void box_edge(int d, double ratio, int srcSize, int &lo, int &hi) {
    lo = (int)__dmul_rn((double)d, ratio);
    hi = (int)__dmul_rn((double)(d + 1), ratio);
    if (hi > srcSize) hi = srcSize;
    if (lo >= hi) lo = hi - 1;
    if (lo < 0) lo = 0;
}

__device__ __forceinline__ uint32_t box_finish(uint32_t sr, uint32_t sg, uint32_t sb, uint32_t sa, int count) {
    double inv = __ddiv_rn(1.0, (double)count);  // ssim.go:302
    uint32_t r = clampf_dev(__dmul_rn((double)sr, inv));
    uint32_t g = clampf_dev(__dmul_rn((double)sg, inv));
    uint32_t b = clampf_dev(__dmul_rn((double)sb, inv));
    uint32_t a = clampf_dev(__dmul_rn((double)sa, inv));
    return r | (g << 8) | (b << 16) | (a << 24);
}

__global__ void __launch_bounds__(kThreads) box_rows_kernel(const BoxParams pIn) {
    __shared__ uint2 colsum[kSpanMax + 8];  // per source column: (R | B<<16, G | A<<16) 16-bit sums
    BoxParams p = pIn;
    int img = blockIdx.z;
    if (img >= p.nFirst) {   // second batch: same dims, its own base / strides
        img -= p.nFirst;
        p.src = p.src2; p.dst = p.dst2; p.srcImgStride = p.src2ImgStride; p.srcRowStride = p.src2RowStride;
    }
    const int dx0 = blockIdx.x * p.dxChunk;
    const int dx1 = min(dx0 + p.dxChunk, p.dstW);
    const int dyEnd = min((int)(blockIdx.y + 1) * p.dyPerCta, p.dstH);
  for (int dy = blockIdx.y * p.dyPerCta; dy < dyEnd; dy++) {
    int sy0, sy1, sxa, sxb, tmp;
    box_edge(dy, p.yRatio, p.srcH, sy0, sy1);
    box_edge(dx0, p.xRatio, p.srcW, sxa, tmp);
    box_edge(dx1 - 1, p.xRatio, p.srcW, tmp, sxb);
    const int base = sxa & ~3;            // 16-byte aligned first column
    const int span = sxb - base;          // <= kSpanMax + 3 by construction of dxChunk
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;

    for (int c4 = threadIdx.x * 4; c4 < span; c4 += kThreads * 4) {
        const int x = base + c4;
        uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
        const bool full = p.vecOK && (x + 4 <= p.srcW);
        const uint8_t *q = s + (long long)sy0 * p.srcRowStride + (long long)x * 4;
        int y = sy0;
        // [An L2 prefetch of the next output row's box was measured slower here: 0.301 against 0.281 ms per 16 pairs of
        // 4032x3024 — 30+ warps per SM already hide the latency (profiles/r2_tuning_sweep.txt).]
        if (full) {  // four independent 128-bit row loads in flight per thread
            for (; y + 3 < sy1; y += 4, q += 4 * (long long)p.srcRowStride) {
                const uint4 t[4] = {ld_nc_u128(q), ld_nc_u128(q + p.srcRowStride), ld_nc_u128(q + 2 * (long long)p.srcRowStride),
                                    ld_nc_u128(q + 3 * (long long)p.srcRowStride)};
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint32_t v[4] = {t[r].x, t[r].y, t[r].z, t[r].w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        lo[i] += v[i] & 0x00FF00FFu;
                        hi[i] += (v[i] >> 8) & 0x00FF00FFu;
                    }
                }
            }
        }
        for (; y < sy1; y++, q += p.srcRowStride) {
            uint32_t v[4];
            if (full) {
                uint4 t = ld_nc_u128(q);
                v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) v[i] = (x + i < p.srcW) ? ld_nc_u32(q + 4 * i) : 0u;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                lo[i] += v[i] & 0x00FF00FFu;          // R, B
                hi[i] += (v[i] >> 8) & 0x00FF00FFu;   // G, A
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) colsum[c4 + i] = make_uint2(lo[i], hi[i]);
    }
    __syncthreads();
    const int count_y = sy1 - sy0;
    uint8_t *drow = p.dst + (long long)img * p.dstImgStride + (long long)dy * p.dstRowStride;
    for (int dx = dx0 + threadIdx.x; dx < dx1; dx += kThreads) {
        int sx0, sx1;
        box_edge(dx, p.xRatio, p.srcW, sx0, sx1);
        uint32_t sr = 0, sg = 0, sb = 0, sa = 0;
        for (int x = sx0; x < sx1; x++) {
            uint2 c = colsum[x - base];
            sr += c.x & 0xFFFFu; sb += c.x >> 16;
            sg += c.y & 0xFFFFu; sa += c.y >> 16;
        }
        *reinterpret_cast<uint32_t *>(drow + (long long)dx * 4) = box_finish(sr, sg, sb, sa, count_y * (sx1 - sx0));
    }
    __syncthreads();  // colsum is reused by the next output row
  }
}

// Exact 2x2 boxes (srcW == 2*dstW, srcH == 2*dstH: the MSSSIM cascade, ssim.go:354-360): the mean of four
// bytes times 0.25 is exact in binary64, so clampF(sum * (1.0/4)) == (sum + 2) >> 2.  Pure streaming:
// each thread reads 4 pixels from two rows (two 128-bit loads) and writes 2 pixels; SIMD on 16-bit lanes.
__global__ void __launch_bounds__(kThreads) box2x_kernel(const BoxParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;   // 4 source px = 2 dest px per thread
    const int dy = blockIdx.y, img = blockIdx.z;
    const int sx = t * 4;
    if (sx >= p.srcW) return;
    const uint8_t *r0 = p.src + (long long)img * p.srcImgStride + (long long)(2 * dy) * p.srcRowStride + (long long)sx * 4;
    const uint8_t *r1 = r0 + p.srcRowStride;
    uint32_t a[4], b[4];
    if (p.vecOK && sx + 4 <= p.srcW) {
        uint4 u = ld_nc_u128(r0), v = ld_nc_u128(r1);
        a[0] = u.x; a[1] = u.y; a[2] = u.z; a[3] = u.w;
        b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            a[i] = (sx + i < p.srcW) ? ld_nc_u32(r0 + 4 * i) : 0u;
            b[i] = (sx + i < p.srcW) ? ld_nc_u32(r1 + 4 * i) : 0u;
        }
    }
    uint32_t out[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const uint32_t rb = (a[2 * k] & 0x00FF00FFu) + (a[2 * k + 1] & 0x00FF00FFu) + (b[2 * k] & 0x00FF00FFu) + (b[2 * k + 1] & 0x00FF00FFu);
        const uint32_t ga = ((a[2 * k] >> 8) & 0x00FF00FFu) + ((a[2 * k + 1] >> 8) & 0x00FF00FFu) +
                            ((b[2 * k] >> 8) & 0x00FF00FFu) + ((b[2 * k + 1] >> 8) & 0x00FF00FFu);
        const uint32_t mrb = ((rb + 0x00020002u) >> 2) & 0x00FF00FFu;
        const uint32_t mga = ((ga + 0x00020002u) >> 2) & 0x00FF00FFu;
        out[k] = mrb | (mga << 8);
    }
    uint8_t *d = p.dst + (long long)img * p.dstImgStride + (long long)dy * p.dstRowStride + (long long)(sx / 2) * 4;
    const int dx = sx / 2;
    if (dx + 2 <= p.dstW && (((uintptr_t)d) & 7) == 0) {
        *reinterpret_cast<uint2 *>(d) = make_uint2(out[0], out[1]);
    } else {
        if (dx < p.dstW) *reinterpret_cast<uint32_t *>(d) = out[0];
        if (dx + 1 < p.dstW) *reinterpret_cast<uint32_t *>(d + 4) = out[1];
    }
}


// ------------------------------------------------------------------------------------------------
// MS-SSIM level kernel: ONE read of a level image produces both things ssim.go needs from it —
// the SSIMFast thumbnail (boxDownsample to <= 512 px, ssim.go:57-58) and the next level's image
// (boxDownsample to w/2 x h/2, ssim.go:354-360) — for both images of the pair batch in one launch.
// Structure of box_rows_kernel; while a thread walks the rows of its thumbnail box it also pairs
// every even row 2k with row 2k+1 (loading one extra row when the box ends on an even row) and
// emits the 2x2 means of its four columns.  Each half-resolution row is produced exactly once, by
// the box that contains its even source row.  Column groups that straddle two chunks are written
// twice with identical bytes.  Preconditions (checked by launch_box_fused): srcW % 4 == 0,
// srcH % 2 == 0, 16-byte aligned rows, thumbnail boxes tile the source exactly.
// ------------------------------------------------------------------------------------------------
struct BoxFusedParams {
    const uint8_t *src[2];
    uint8_t *thumb[2];
    uint8_t *half[2];
    long long srcImgStride[2];
    int srcRowStride[2];
    long long thumbImgStride, halfImgStride;
    int thumbRowStride, halfRowStride;
    int srcW, srcH, dstW, dstH, n;
    double xRatio, yRatio;
    int dxChunk, dyPerCta;
};

__device__ __forceinline__ uint2 mean2x2(const uint32_t (&a)[4], const uint32_t (&b)[4]) {
    uint32_t out[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const uint32_t rb = (a[2 * k] & 0x00FF00FFu) + (a[2 * k + 1] & 0x00FF00FFu) + (b[2 * k] & 0x00FF00FFu) + (b[2 * k + 1] & 0x00FF00FFu);
        const uint32_t ga = ((a[2 * k] >> 8) & 0x00FF00FFu) + ((a[2 * k + 1] >> 8) & 0x00FF00FFu) +
                            ((b[2 * k] >> 8) & 0x00FF00FFu) + ((b[2 * k + 1] >> 8) & 0x00FF00FFu);
        out[k] = (((rb + 0x00020002u) >> 2) & 0x00FF00FFu) | ((((ga + 0x00020002u) >> 2) & 0x00FF00FFu) << 8);
    }
    return make_uint2(out[0], out[1]);
}

__global__ void __launch_bounds__(kThreads) box_fused_kernel(const BoxFusedParams p) {
    __shared__ uint2 colsum[kSpanMax + 8];
    const int which = (int)blockIdx.z >= p.n ? 1 : 0;
    const int img = (int)blockIdx.z - which * p.n;
    const int rs = p.srcRowStride[which];
    const uint8_t *s = p.src[which] + (long long)img * p.srcImgStride[which];
    uint8_t *hd = p.half[which] + (long long)img * p.halfImgStride;
    uint8_t *td = p.thumb[which] + (long long)img * p.thumbImgStride;
    const int dx0 = blockIdx.x * p.dxChunk;
    const int dx1 = min(dx0 + p.dxChunk, p.dstW);
    const int dyEnd = min((int)(blockIdx.y + 1) * p.dyPerCta, p.dstH);
    int sxa, sxb, tmp;
    box_edge(dx0, p.xRatio, p.srcW, sxa, tmp);
    box_edge(dx1 - 1, p.xRatio, p.srcW, tmp, sxb);
    const int base = sxa & ~3;
    const int span = sxb - base;
    for (int dy = blockIdx.y * p.dyPerCta; dy < dyEnd; dy++) {
        int sy0, sy1;
        box_edge(dy, p.yRatio, p.srcH, sy0, sy1);
        for (int c4 = threadIdx.x * 4; c4 < span; c4 += kThreads * 4) {
            const int x = base + c4;   // x + 4 <= srcW (srcW % 4 == 0 and x < sxb <= srcW)
            uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
            const uint8_t *q = s + (long long)sy0 * rs + (long long)x * 4;
            uint8_t *hcol = hd + (long long)(x >> 1) * 4;
            auto acc = [&](const uint4 &t) {
                const uint32_t v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    lo[i] += v[i] & 0x00FF00FFu;          // R, B
                    hi[i] += (v[i] >> 8) & 0x00FF00FFu;   // G, A
                }
            };
            auto emit = [&](const uint4 &ta, const uint4 &tb, int hy) {
                const uint32_t a[4] = {ta.x, ta.y, ta.z, ta.w}, b[4] = {tb.x, tb.y, tb.z, tb.w};
                *reinterpret_cast<uint2 *>(hcol + (long long)hy * p.halfRowStride) = mean2x2(a, b);
            };
            int y = sy0;
            if (y & 1) {  // the box starts on an odd row: its partner (and the output row) belong to the previous box
                acc(ld_nc_u128(q));
                q += rs;
                y++;
            }
            // row pairs (2k, 2k+1) inside the box: two pairs (four independent 128-bit loads) in flight per thread
            for (; y + 3 < sy1; y += 4, q += 4 * (long long)rs) {
                const uint4 t0 = ld_nc_u128(q), t1 = ld_nc_u128(q + rs), t2 = ld_nc_u128(q + 2 * (long long)rs),
                            t3 = ld_nc_u128(q + 3 * (long long)rs);
                acc(t0); acc(t1); acc(t2); acc(t3);
                emit(t0, t1, y >> 1);
                emit(t2, t3, (y >> 1) + 1);
            }
            if (y + 1 < sy1) {
                const uint4 t0 = ld_nc_u128(q), t1 = ld_nc_u128(q + rs);
                acc(t0); acc(t1);
                emit(t0, t1, y >> 1);
                y += 2;
                q += 2 * (long long)rs;
            }
            if (y < sy1) {  // the box ends on an even row: fetch its partner (first row of the next box; srcH is even)
                const uint4 t0 = ld_nc_u128(q), t1 = ld_nc_u128(q + rs);
                acc(t0);
                emit(t0, t1, y >> 1);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) colsum[c4 + i] = make_uint2(lo[i], hi[i]);
        }
        __syncthreads();
        const int count_y = sy1 - sy0;
        uint8_t *drow = td + (long long)dy * p.thumbRowStride;
        for (int dx = dx0 + threadIdx.x; dx < dx1; dx += kThreads) {
            int sx0, sx1;
            box_edge(dx, p.xRatio, p.srcW, sx0, sx1);
            uint32_t sr = 0, sg = 0, sb = 0, sa = 0;
            for (int xx = sx0; xx < sx1; xx++) {
                uint2 c = colsum[xx - base];
                sr += c.x & 0xFFFFu; sb += c.x >> 16;
                sg += c.y & 0xFFFFu; sa += c.y >> 16;
            }
            *reinterpret_cast<uint32_t *>(drow + (long long)dx * 4) = box_finish(sr, sg, sb, sa, count_y * (sx1 - sx0));
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Two MS-SSIM levels from ONE read ("level 1 is re-read from HBM instead of being produced while level 0's
// tile is on chip").  A CTA owns a band of level-l rows x a chunk of level-l columns chosen on the host so that every
// product has whole rows / columns inside it: band and chunk boundaries are box edges of BOTH thumbnails, multiples of
// 4 rows and 8 columns.  The level-(l+1) pixels — the rounded bytes boxDownsample would write (ssim.go:354-360,
// 286-309) — only ever exist in registers: a thread turns 4 rows x 8 columns of level l into 2 x 4 pixels of level
// l+1 and 1 x 2 pixels of level l+2, and feeds both levels' rows into the column sums of their thumbnails.  8K pairs
// move 133 + 8.3 + 1.2 MB per image instead of 133 + 33 + 33 + 8.3 + 1.2.
// (First version, measured slower than the two single-level passes: a 28 KB shared-memory tile for level l+1 and a
// second walk over it — 2 CTAs per SM, 1.78 ms against 1.47 ms per 16 8K pairs.)
// ------------------------------------------------------------------------------------------------
constexpr int kF2Threads = 256;
constexpr int kF2MaxBand = 64;        // level-l rows per band
constexpr int kF2MaxChunk = 2048;     // level-l columns per chunk
constexpr int kF2SmemBudget = 96 * 1024;
struct BoxFused2Params {
    const uint8_t *src[2];
    long long srcImgStride[2];
    int srcRowStride[2];
    uint8_t *thumb0[2], *thumb1[2], *l2[2];
    long long thumb0ImgStride, thumb1ImgStride, l2ImgStride;
    int thumb0RowStride, thumb1RowStride, l2RowStride;
    int srcW, srcH, n;
    int tw0, th0, tw1, th1;
    double xr0, yr0, xr1, yr1;
    int bandRows, chunkCols;          // level-l rows per band, columns per chunk
    int t0PerBand, t0PerChunk;        // level-l thumbnail rows per band / columns per chunk
    int t1PerBand, t1PerChunk;        // level-(l+1) thumbnail rows per band / columns per chunk
};

// Packed 16-bit channel sums of one thumbnail row, flushed to shared memory when the walk crosses a box edge (runs of
// different threads may end inside the same box, hence atomics).
template <int NCOL>
struct F2Acc {
    uint32_t lo[NCOL], hi[NCOL];
    int ty, tyEnd;
    __device__ __forceinline__ void init(int dyFirst, double ratio, int imgRows, int firstRow) {
#pragma unroll
        for (int i = 0; i < NCOL; i++) lo[i] = hi[i] = 0;
        int lo_;
        ty = dyFirst;
        box_edge(ty, ratio, imgRows, lo_, tyEnd);
        while (tyEnd <= firstRow) { ty++; box_edge(ty, ratio, imgRows, lo_, tyEnd); }
    }
    __device__ __forceinline__ void flush(uint2 *colsum, int dyFirst, int pitch, int col0) {
        uint2 *cs = colsum + (size_t)(ty - dyFirst) * pitch + col0;
#pragma unroll
        for (int i = 0; i < NCOL; i++) {
            atomicAdd(&cs[i].x, lo[i]);
            atomicAdd(&cs[i].y, hi[i]);
            lo[i] = hi[i] = 0;
        }
    }
    // one image row (NCOL pixels) at row y
    __device__ __forceinline__ void add(const uint32_t (&v)[NCOL], int y, uint2 *colsum, int dyFirst, int pitch, int col0, double ratio,
                                        int imgRows) {
        if (y >= tyEnd) {   // crossed into the next thumbnail row
            flush(colsum, dyFirst, pitch, col0);
            ty++;
            int lo_;
            box_edge(ty, ratio, imgRows, lo_, tyEnd);
        }
#pragma unroll
        for (int i = 0; i < NCOL; i++) {
            lo[i] += v[i] & 0x00FF00FFu;          // R, B
            hi[i] += (v[i] >> 8) & 0x00FF00FFu;   // G, A
        }
    }
};

// Thumbnail pixels of `nTy` thumbnail rows x [dx0, dx1) from the column sums (one thread per output pixel).
__device__ __forceinline__ void f2_outputs(const uint2 *colsum, int csPitch, int x0, int dyFirst, int nTy, int dx0, int dx1,
                                           double xRatio, double yRatio, int imgW, int imgH, uint8_t *td, int tdRowStride) {
    const int nTx = dx1 - dx0;
    for (int idx = threadIdx.x; idx < nTy * nTx; idx += kF2Threads) {
        const int ty = idx / nTx, dx = dx0 + idx - ty * nTx;
        int sy0, sy1, sx0, sx1;
        box_edge(dyFirst + ty, yRatio, imgH, sy0, sy1);
        box_edge(dx, xRatio, imgW, sx0, sx1);
        const uint2 *cs = colsum + (size_t)ty * csPitch - x0;
        uint32_t sr = 0, sg = 0, sb = 0, sa = 0;
        for (int xx = sx0; xx < sx1; xx++) {
            const uint2 c = cs[xx];
            sr += c.x & 0xFFFFu; sb += c.x >> 16;
            sg += c.y & 0xFFFFu; sa += c.y >> 16;
        }
        *reinterpret_cast<uint32_t *>(td + (long long)(dyFirst + ty) * tdRowStride + (long long)dx * 4) =
            box_finish(sr, sg, sb, sa, (sy1 - sy0) * (sx1 - sx0));
    }
}

// A thread owns a stripe of 4 level-l columns (one 128-bit load per row, contiguous across the warp) and a run of
// 4-row steps.  One step = 4 loads -> 2 rows x 2 pixels of level l+1 (registers only) -> 1 pixel of level l+2 (a 4-byte
// store, contiguous across the warp); two steps are in flight at a time.  Every level-l row and every level-(l+1) row
// is added to the column sums of the thumbnail row that contains it.
__device__ __forceinline__ uint32_t mean2x2_px(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    const uint32_t rb = (a0 & 0x00FF00FFu) + (a1 & 0x00FF00FFu) + (b0 & 0x00FF00FFu) + (b1 & 0x00FF00FFu);
    const uint32_t ga = ((a0 >> 8) & 0x00FF00FFu) + ((a1 >> 8) & 0x00FF00FFu) + ((b0 >> 8) & 0x00FF00FFu) + ((b1 >> 8) & 0x00FF00FFu);
    return (((rb + 0x00020002u) >> 2) & 0x00FF00FFu) | ((((ga + 0x00020002u) >> 2) & 0x00FF00FFu) << 8);
}

#ifndef FB_F2_SINGLE
#define FB_F2_SINGLE 1
#endif
#ifndef FB_F2_SPF
#define FB_F2_SPF 2   // single-step loop: L2 prefetch distance in steps
#endif
#ifndef FB_F2_PF
#define FB_F2_PF 1
#endif
template <int MINB>
__global__ void __launch_bounds__(kF2Threads, MINB) box_fused2_kernel(const BoxFused2Params p) {
    extern __shared__ __align__(16) uint8_t f2smem[];
    const int which = (int)blockIdx.z >= p.n ? 1 : 0;
    const int img = (int)blockIdx.z - which * p.n;
    const int rs = p.srcRowStride[which];
    const uint8_t *s = p.src[which] + (long long)img * p.srcImgStride[which];
    uint8_t *l2 = p.l2[which] + (long long)img * p.l2ImgStride;
    const int X0 = blockIdx.x * p.chunkCols, Y0 = blockIdx.y * p.bandRows;
    const int X1 = min(X0 + p.chunkCols, p.srcW), Y1 = min(Y0 + p.bandRows, p.srcH);
    const int cols = X1 - X0, rows = Y1 - Y0;                  // multiples of 8 and 4
    const int h1 = p.srcH >> 1;
    const int dy0 = blockIdx.y * p.t0PerBand, nTy0 = min(p.t0PerBand, p.th0 - dy0);
    const int dy1 = blockIdx.y * p.t1PerBand, nTy1 = min(p.t1PerBand, p.th1 - dy1);
    uint2 *colsum0 = reinterpret_cast<uint2 *>(f2smem);                       // [t0PerBand][chunkCols]
    uint2 *colsum1 = colsum0 + (size_t)p.t0PerBand * p.chunkCols;              // [t1PerBand][chunkCols / 2]
    const int pitch0 = p.chunkCols, pitch1 = p.chunkCols >> 1;
    for (int i = threadIdx.x; i < p.t0PerBand * pitch0 + p.t1PerBand * pitch1; i += kF2Threads) colsum0[i] = make_uint2(0u, 0u);
    __syncthreads();
    {
        const int stripes = cols >> 2, steps = rows >> 2;
        int groups = kF2Threads / stripes;
        if (groups < 1) groups = 1;
        if (groups > steps) groups = steps;
        const int spg = (steps + groups - 1) / groups;
        for (int work = threadIdx.x; work < stripes * groups; work += kF2Threads) {
            const int st = work % stripes, g = work / stripes;
            const int k0 = g * spg, k1 = min(k0 + spg, steps);
            if (k0 >= k1) continue;
            F2Acc<4> a0;
            F2Acc<2> a1;
            a0.init(dy0, p.yr0, p.srcH, Y0 + 4 * k0);
            a1.init(dy1, p.yr1, h1, (Y0 >> 1) + 2 * k0);
            const uint8_t *q = s + (long long)(Y0 + 4 * k0) * rs + (long long)(X0 + 4 * st) * 4;
            uint8_t *o = l2 + (long long)((Y0 >> 2) + k0) * p.l2RowStride + (long long)((X0 >> 2) + st) * 4;
            auto step = [&](const uint4 (&t)[4], int k, uint8_t *op) {
                uint32_t L1[2][2];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint32_t v[4] = {t[r].x, t[r].y, t[r].z, t[r].w};
                    a0.add(v, Y0 + 4 * k + r, colsum0, dy0, pitch0, 4 * st, p.yr0, p.srcH);
                }
#pragma unroll
                for (int hr = 0; hr < 2; hr++) {
                    L1[hr][0] = mean2x2_px(t[2 * hr].x, t[2 * hr].y, t[2 * hr + 1].x, t[2 * hr + 1].y);
                    L1[hr][1] = mean2x2_px(t[2 * hr].z, t[2 * hr].w, t[2 * hr + 1].z, t[2 * hr + 1].w);
                    a1.add(L1[hr], (Y0 >> 1) + 2 * k + hr, colsum1, dy1, pitch1, 2 * st, p.yr1, h1);
                }
                *reinterpret_cast<uint32_t *>(op) = mean2x2_px(L1[0][0], L1[0][1], L1[1][0], L1[1][1]);
            };
            // [Measured and dropped (profiles/r2_tuning_sweep.txt): register double-buffering of the NEXT pair of steps (needs 2
            // CTAs per SM: 1.16 ms per 16 8K pairs against 1.13) and a rolling two-step pipeline at the same register
            // count (ptxas spills at 80 registers: 1.44-1.60 ms; 1.23 at 2 CTAs).  The L2 prefetch below costs nothing.]
            int k = k0;
#if FB_F2_SINGLE   // one step (four loads) per iteration with an L2 prefetch FB_F2_SPF steps ahead: 64 registers, so a 4th CTA fits the SM
                   // (16 8K pairs: 0.960 ms against 1.064 ms with two steps in flight at 3 CTAs; prefetch 1 / 3 / 4 / 6 steps ahead:
                   // 0.995 / 0.973 / 0.971 / 1.001; 3 or 5 CTAs: 1.021 / 1.217 — profiles/r2_tuning_sweep.txt)
            for (; k < k1; k++, q += 4 * (long long)rs, o += (long long)p.l2RowStride) {
                uint4 ta[4];
                if (k + FB_F2_SPF + 1 <= k1) {
#pragma unroll
                    for (int r = 0; r < 4; r++) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + (long long)(4 * FB_F2_SPF + r) * rs));
                }
#pragma unroll
                for (int r = 0; r < 4; r++) ta[r] = ld_nc_u128(q + (long long)r * rs);
                step(ta, k, o);
            }
#endif
            for (; k + 2 <= k1; k += 2, q += 8 * (long long)rs, o += 2 * (long long)p.l2RowStride) {
                uint4 ta[4], tb[4];
#if FB_F2_PF > 0
                // L2 prefetch of this thread's 16 bytes of the rows FB_F2_PF iterations ahead (inside the band): the loads
                // below then find them in L2 (~300 cycles) instead of DRAM; no registers, 8 instructions per 8 rows.
                if (k + 2 * FB_F2_PF + 2 <= k1) {
#pragma unroll
                    for (int r = 0; r < 8; r++)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + (long long)(8 * FB_F2_PF + r) * rs));
                }
#endif
#pragma unroll
                for (int r = 0; r < 4; r++) ta[r] = ld_nc_u128(q + (long long)r * rs);
#pragma unroll
                for (int r = 0; r < 4; r++) tb[r] = ld_nc_u128(q + (long long)(4 + r) * rs);
                step(ta, k, o);
                step(tb, k + 1, o + p.l2RowStride);
            }
            if (k < k1) {
                uint4 ta[4];
#pragma unroll
                for (int r = 0; r < 4; r++) ta[r] = ld_nc_u128(q + (long long)r * rs);
                step(ta, k, o);
            }
            a0.flush(colsum0, dy0, pitch0, 4 * st);
            a1.flush(colsum1, dy1, pitch1, 2 * st);
        }
    }
    __syncthreads();
    f2_outputs(colsum0, pitch0, X0, dy0, nTy0, blockIdx.x * p.t0PerChunk, min((int)(blockIdx.x + 1) * p.t0PerChunk, p.tw0), p.xr0, p.yr0,
               p.srcW, p.srcH, p.thumb0[which] + (long long)img * p.thumb0ImgStride, p.thumb0RowStride);
    f2_outputs(colsum1, pitch1, X0 >> 1, dy1, nTy1, blockIdx.x * p.t1PerChunk, min((int)(blockIdx.x + 1) * p.t1PerChunk, p.tw1), p.xr1,
               p.yr1, p.srcW >> 1, h1, p.thumb1[which] + (long long)img * p.thumb1ImgStride, p.thumb1RowStride);
}

// Generic fallback: one thread per output pixel walks its own box (upsampling, boxes taller than
// 256 rows or wider than the shared-memory span). Same arithmetic.
__global__ void __launch_bounds__(kThreads) box_naive_kernel(const BoxParams p) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    const int dy = blockIdx.y, img = blockIdx.z;
    if (dx >= p.dstW) return;
    int sy0, sy1, sx0, sx1;
    box_edge(dy, p.yRatio, p.srcH, sy0, sy1);
    box_edge(dx, p.xRatio, p.srcW, sx0, sx1);
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;
    unsigned long long sr = 0, sg = 0, sb = 0, sa = 0;
    for (int y = sy0; y < sy1; y++) {
        const uint8_t *q = s + (long long)y * p.srcRowStride + (long long)sx0 * 4;
        for (int x = sx0; x < sx1; x++, q += 4) {
            uint32_t v = ld_nc_u32(q);
            sr += v & 0xFF; sg += (v >> 8) & 0xFF; sb += (v >> 16) & 0xFF; sa += v >> 24;
        }
    }
    long long count = (long long)(sy1 - sy0) * (sx1 - sx0);
    uint8_t *d = p.dst + (long long)img * p.dstImgStride + (long long)dy * p.dstRowStride + (long long)dx * 4;
    uint32_t out = 0u;  // an empty box leaves the (zero-filled) destination pixel untouched (ssim.go:301)
    if (count > 0) {
        double inv = __ddiv_rn(1.0, (double)count);
        uint32_t r = clampf_dev(__dmul_rn((double)sr, inv));
        uint32_t g = clampf_dev(__dmul_rn((double)sg, inv));
        uint32_t b = clampf_dev(__dmul_rn((double)sb, inv));
        uint32_t a = clampf_dev(__dmul_rn((double)sa, inv));
        out = r | (g << 8) | (b << 16) | (a << 24);
    }
    *reinterpret_cast<uint32_t *>(d) = out;
}

}  // namespace

// Host copy of the edge rule (used to size chunks; same IEEE ops as the device).
void box_edges_host(int src, int dst, int *lo, int *hi) {
    double ratio = (double)src / (double)dst;
    for (int d = 0; d < dst; d++) {
        int a = (int)((double)d * ratio);
        int b = (int)((double)(d + 1) * ratio);
        if (b > src) b = src;
        if (a >= b) a = b - 1;
        if (a < 0) a = 0;
        lo[d] = a;
        hi[d] = b;
    }
}

static int launch_box_impl(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int srcW,
                           int srcH, uint8_t *dst, long long dstImgStride, int dstRowStride, int dstW, int dstH, int n,
                           const uint8_t *src2, long long src2ImgStride, int src2RowStride, uint8_t *dst2);

int launch_box(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int srcW,
               int srcH, uint8_t *dst, long long dstImgStride, int dstRowStride, int dstW, int dstH, int n,
               const int *unused_edges) {
    (void)unused_edges;
    return launch_box_impl(s, src, srcImgStride, srcRowStride, srcW, srcH, dst, dstImgStride, dstRowStride, dstW, dstH, n,
                           nullptr, 0, 0, nullptr);
}

// Both images of every pair in one launch (same dims; the destinations share strides): one grid instead of two
// half-sized ones, so the tail of the first does not idle the GPU.
int launch_box_pair(cudaStream_t s, const uint8_t *srcA, long long srcImgStrideA, int srcRowStrideA, const uint8_t *srcB,
                    long long srcImgStrideB, int srcRowStrideB, int srcW, int srcH, uint8_t *dstA, uint8_t *dstB,
                    long long dstImgStride, int dstRowStride, int dstW, int dstH, int n) {
    return launch_box_impl(s, srcA, srcImgStrideA, srcRowStrideA, srcW, srcH, dstA, dstImgStride, dstRowStride, dstW, dstH, n,
                           srcB, srcImgStrideB, srcRowStrideB, dstB);
}

static int launch_box_impl(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int srcW,
                           int srcH, uint8_t *dst, long long dstImgStride, int dstRowStride, int dstW, int dstH, int n,
                           const uint8_t *src2, long long src2ImgStride, int src2RowStride, uint8_t *dst2) {
    if (n <= 0) return FB_OK;
    BoxParams p;
    p.src2 = src2; p.dst2 = dst2; p.src2ImgStride = src2ImgStride; p.src2RowStride = src2RowStride; p.nFirst = n;
    p.src = src; p.dst = dst;
    p.srcImgStride = srcImgStride; p.dstImgStride = dstImgStride;
    p.srcRowStride = srcRowStride; p.dstRowStride = dstRowStride;
    p.srcW = srcW; p.srcH = srcH; p.dstW = dstW; p.dstH = dstH;
    p.xRatio = (double)srcW / (double)dstW;   // ssim.go:251-252
    p.yRatio = (double)srcH / (double)dstH;
    p.vecOK = (((uintptr_t)src | (uintptr_t)srcImgStride | (uintptr_t)srcRowStride | (uintptr_t)src2 | (uintptr_t)src2ImgStride |
                 (uintptr_t)src2RowStride) & 15) == 0;
    // Fast path needs: disjoint ascending boxes (ratio >= 1), <= 256 rows per box (16-bit sums),
    // and at least one output column per shared-memory span.
    int maxBoxW = (int)p.xRatio + 2, maxBoxH = (int)p.yRatio + 2;
    bool fast = p.xRatio >= 1.0 && p.yRatio >= 1.0 && maxBoxH <= 256 && maxBoxW <= kSpanMax / 2;
    const bool exact2x = srcW == 2 * dstW && srcH == 2 * dstH;
    if (src2 != nullptr && (exact2x || !fast)) {   // only box_rows_kernel understands the second batch
        int rc = launch_box_impl(s, src, srcImgStride, srcRowStride, srcW, srcH, dst, dstImgStride, dstRowStride, dstW, dstH, n, nullptr, 0, 0, nullptr);
        if (rc != FB_OK) return rc;
        return launch_box_impl(s, src2, src2ImgStride, src2RowStride, srcW, srcH, dst2, dstImgStride, dstRowStride, dstW, dstH, n, nullptr, 0, 0, nullptr);
    }
    if (exact2x) {  // every box is exactly 2x2
        p.dxChunk = 0; p.dyPerCta = 1;
        dim3 grid(((srcW + 3) / 4 + kThreads - 1) / kThreads, dstH, n);
        box2x_kernel<<<grid, kThreads, 0, s>>>(p);
    } else if (fast) {
        int chunk = (int)((double)(kSpanMax - maxBoxW - 4) / p.xRatio);
        if (chunk < 1) chunk = 1;
        if (chunk > dstW) chunk = dstW;
        p.dxChunk = chunk;
        // ~32 source rows per CTA: one output row for 15x15 boxes, 16 output rows for the 2x cascade
        p.dyPerCta = maxBoxH >= 32 ? 1 : (32 / maxBoxH < 1 ? 1 : 32 / maxBoxH);
        dim3 grid((dstW + chunk - 1) / chunk, (dstH + p.dyPerCta - 1) / p.dyPerCta, src2 ? 2 * n : n);
        box_rows_kernel<<<grid, kThreads, 0, s>>>(p);
    } else {
        p.dxChunk = 0;
        p.dyPerCta = 1;
        dim3 grid((dstW + kThreads - 1) / kThreads, dstH, n);
        box_naive_kernel<<<grid, kThreads, 0, s>>>(p);
    }
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

// Boxes [lo, hi) of a `src -> dst` downsample tile [0, src) exactly (contiguous, complete)?
static bool boxes_tile(int src, int dst) {
    std::vector<int> lo(dst), hi(dst);
    box_edges_host(src, dst, lo.data(), hi.data());
    if (lo[0] != 0 || hi[dst - 1] != src) return false;
    for (int d = 0; d + 1 < dst; d++)
        if (hi[d] != lo[d + 1] || hi[d] <= lo[d]) return false;
    return true;
}

// Thumbnail (tw x th) + half-resolution image of both batches from one read.  Returns 1 (nothing launched)
// when the preconditions of box_fused_kernel do not hold; the caller then uses launch_box twice per image.
int launch_box_fused(cudaStream_t s, const uint8_t *srcA, long long srcImgStrideA, int srcRowStrideA,
                     const uint8_t *srcB, long long srcImgStrideB, int srcRowStrideB, int srcW, int srcH,
                     uint8_t *thumbA, uint8_t *thumbB, long long thumbImgStride, int thumbRowStride, int tw, int th,
                     uint8_t *halfA, uint8_t *halfB, long long halfImgStride, int halfRowStride, int n) {
    if (n <= 0) return FB_OK;
    if (getenv("FB_BOX_NOFUSE") != nullptr) return 1;
    if ((srcW & 3) || (srcH & 1) || tw < 1 || th < 1) return 1;
    const uintptr_t al = (uintptr_t)srcA | (uintptr_t)srcB | (uintptr_t)srcImgStrideA | (uintptr_t)srcImgStrideB |
                         (uintptr_t)srcRowStrideA | (uintptr_t)srcRowStrideB;
    if ((al & 15) || (((uintptr_t)halfA | (uintptr_t)halfB | (uintptr_t)halfImgStride | (uintptr_t)halfRowStride) & 7)) return 1;
    BoxFusedParams p;
    p.src[0] = srcA; p.src[1] = srcB;
    p.srcImgStride[0] = srcImgStrideA; p.srcImgStride[1] = srcImgStrideB;
    p.srcRowStride[0] = srcRowStrideA; p.srcRowStride[1] = srcRowStrideB;
    p.thumb[0] = thumbA; p.thumb[1] = thumbB; p.thumbImgStride = thumbImgStride; p.thumbRowStride = thumbRowStride;
    p.half[0] = halfA; p.half[1] = halfB; p.halfImgStride = halfImgStride; p.halfRowStride = halfRowStride;
    p.srcW = srcW; p.srcH = srcH; p.dstW = tw; p.dstH = th; p.n = n;
    p.xRatio = (double)srcW / (double)tw;
    p.yRatio = (double)srcH / (double)th;
    const int maxBoxW = (int)p.xRatio + 2, maxBoxH = (int)p.yRatio + 2;
    if (!(p.xRatio >= 1.0 && p.yRatio >= 1.0 && maxBoxH <= 256 && maxBoxW <= kSpanMax / 2)) return 1;
    if (!boxes_tile(srcW, tw) || !boxes_tile(srcH, th)) return 1;
    int chunk = (int)((double)(kSpanMax - maxBoxW - 4) / p.xRatio);
    if (chunk < 1) chunk = 1;
    if (chunk > tw) chunk = tw;
    p.dxChunk = chunk;
    p.dyPerCta = maxBoxH >= 32 ? 1 : (32 / maxBoxH < 1 ? 1 : 32 / maxBoxH);
    dim3 grid((tw + chunk - 1) / chunk, (th + p.dyPerCta - 1) / p.dyPerCta, 2 * n);
    box_fused_kernel<<<grid, kThreads, 0, s>>>(p);
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

// Smallest period P (multiple of `mult`, <= maxP) such that every multiple of P below `src` is a box edge of the
// src -> dst0 downsample AND (halved) of the src/2 -> dst1 downsample; 0 if none.  *per0 / *per1 = thumbnail boxes per period.
static int common_period(int src, int dst0, int dst1, int mult, int maxP, int *per0, int *per1) {
    std::vector<int> lo0(dst0), hi0(dst0), lo1(dst1), hi1(dst1);
    box_edges_host(src, dst0, lo0.data(), hi0.data());
    box_edges_host(src / 2, dst1, lo1.data(), hi1.data());
    for (int P = mult; P <= maxP; P += mult) {
        if (src % P) continue;
        if ((long long)dst0 * P % src || (long long)dst1 * P % src) continue;
        const int n0 = (int)((long long)dst0 * P / src), n1 = (int)((long long)dst1 * P / src);
        if (n0 < 1 || n1 < 1) continue;
        bool ok = true;
        for (int k = 0; ok && k * P < src; k++) {
            if (lo0[k * n0] != k * P || lo1[k * n1] != k * P / 2) ok = false;
        }
        if (ok) { *per0 = n0; *per1 = n1; return P; }
    }
    return 0;
}

// Levels l and l+1 of the MS-SSIM pyramid from one read of level l: thumbnails of both levels and the level-(l+2)
// image.  Returns 1 (nothing launched) when the geometry has no common period; the caller then takes two single steps.
int launch_box_fused2(cudaStream_t s, const uint8_t *srcA, long long srcImgStrideA, int srcRowStrideA, const uint8_t *srcB,
                      long long srcImgStrideB, int srcRowStrideB, int srcW, int srcH, uint8_t *thumb0A, uint8_t *thumb0B,
                      long long thumb0ImgStride, int thumb0RowStride, int tw0, int th0, uint8_t *thumb1A, uint8_t *thumb1B,
                      long long thumb1ImgStride, int thumb1RowStride, int tw1, int th1, uint8_t *l2A, uint8_t *l2B,
                      long long l2ImgStride, int l2RowStride, int n) {
    if (n <= 0) return FB_OK;
    if (getenv("FB_BOX_NOFUSE") != nullptr || getenv("FB_BOX_NOFUSE2") != nullptr) return 1;
    if ((srcW & 7) || (srcH & 3) || tw0 < 1 || th0 < 1 || tw1 < 1 || th1 < 1) return 1;
    const uintptr_t al = (uintptr_t)srcA | (uintptr_t)srcB | (uintptr_t)srcImgStrideA | (uintptr_t)srcImgStrideB |
                         (uintptr_t)srcRowStrideA | (uintptr_t)srcRowStrideB;
    if ((al & 15) || (((uintptr_t)l2A | (uintptr_t)l2B | (uintptr_t)l2ImgStride | (uintptr_t)l2RowStride) & 7)) return 1;
    BoxFused2Params p;
    p.xr0 = (double)srcW / (double)tw0; p.yr0 = (double)srcH / (double)th0;
    p.xr1 = (double)(srcW / 2) / (double)tw1; p.yr1 = (double)(srcH / 2) / (double)th1;
    auto fits = [](double r) { return r >= 1.0 && (int)r + 2 <= 256; };
    if (!fits(p.xr0) || !fits(p.yr0) || !fits(p.xr1) || !fits(p.yr1)) return 1;
    if (!boxes_tile(srcW, tw0) || !boxes_tile(srcH, th0) || !boxes_tile(srcW / 2, tw1) || !boxes_tile(srcH / 2, th1)) return 1;
    int t0b, t1b, t0c, t1c;
    const int band = common_period(srcH, th0, th1, 4, kF2MaxBand, &t0b, &t1b);
    const int colP = common_period(srcW, tw0, tw1, 8, kF2MaxChunk, &t0c, &t1c);
    if (!band || !colP) return 1;
    // whole periods per CTA: up to 64 rows x ~1024 columns (one 4-column stripe per thread, each thread walking the whole
    // band: 16 8K pairs take 1.60 / 1.28 / 1.12 / 1.23 ms at 240 / 480 / 960 / 1440 columns), shrunk until the column sums
    // of every thumbnail row of the band fit the shared-memory budget
    const int rrep = kF2MaxBand / band;
    static const int chunkTarget = [] { const char *e = getenv("FB_F2_CHUNK"); int v = e ? atoi(e) : 0; return (v >= 64 && v <= kF2MaxChunk) ? v : 1024; }();
    int reps = chunkTarget / colP > 0 ? chunkTarget / colP : 1;
    auto smem_for = [&](int reps_) {
        const size_t chunk = (size_t)colP * reps_;
        return (size_t)t0b * rrep * chunk * sizeof(uint2) + (size_t)t1b * rrep * (chunk / 2) * sizeof(uint2);
    };
    while (reps > 1 && smem_for(reps) > (size_t)kF2SmemBudget) reps--;
    if (smem_for(reps) > (size_t)kF2SmemBudget) return 1;
    p.bandRows = band * rrep; p.chunkCols = colP * reps;
    p.t0PerBand = t0b * rrep; p.t1PerBand = t1b * rrep; p.t0PerChunk = t0c * reps; p.t1PerChunk = t1c * reps;
    p.src[0] = srcA; p.src[1] = srcB;
    p.srcImgStride[0] = srcImgStrideA; p.srcImgStride[1] = srcImgStrideB;
    p.srcRowStride[0] = srcRowStrideA; p.srcRowStride[1] = srcRowStrideB;
    p.thumb0[0] = thumb0A; p.thumb0[1] = thumb0B; p.thumb0ImgStride = thumb0ImgStride; p.thumb0RowStride = thumb0RowStride;
    p.thumb1[0] = thumb1A; p.thumb1[1] = thumb1B; p.thumb1ImgStride = thumb1ImgStride; p.thumb1RowStride = thumb1RowStride;
    p.l2[0] = l2A; p.l2[1] = l2B; p.l2ImgStride = l2ImgStride; p.l2RowStride = l2RowStride;
    p.srcW = srcW; p.srcH = srcH; p.n = n;
    p.tw0 = tw0; p.th0 = th0; p.tw1 = tw1; p.th1 = th1;
    const size_t smem = smem_for(reps);
    static bool attrSet = false;
    if (!attrSet) {
        FB_CUDA(cudaFuncSetAttribute(box_fused2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kF2SmemBudget));
        FB_CUDA(cudaFuncSetAttribute(box_fused2_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        FB_CUDA(cudaFuncSetAttribute(box_fused2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kF2SmemBudget));
        FB_CUDA(cudaFuncSetAttribute(box_fused2_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        FB_CUDA(cudaFuncSetAttribute(box_fused2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kF2SmemBudget));
        FB_CUDA(cudaFuncSetAttribute(box_fused2_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        FB_CUDA(cudaFuncSetAttribute(box_fused2_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, kF2SmemBudget));
        FB_CUDA(cudaFuncSetAttribute(box_fused2_kernel<5>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attrSet = true;
    }
    dim3 grid((srcW + p.chunkCols - 1) / p.chunkCols, (srcH + p.bandRows - 1) / p.bandRows, 2 * n);
    static const int minb = [] { const char *e = getenv("FB_F2_MINB"); return (e && e[0] >= '2' && e[0] <= '5') ? e[0] - '0' : (FB_F2_SINGLE ? 4 : 3); }();
    if (minb == 2) box_fused2_kernel<2><<<grid, kF2Threads, smem, s>>>(p);
    else if (minb == 4) box_fused2_kernel<4><<<grid, kF2Threads, smem, s>>>(p);
    else if (minb == 5) box_fused2_kernel<5><<<grid, kF2Threads, smem, s>>>(p);
    else box_fused2_kernel<3><<<grid, kF2Threads, smem, s>>>(p);
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

}  // namespace fb
