// effects.cu — K5/K6: GaussianBlur, gaussianBlur3x3, Sharpen, AdaptiveSharpen (effects.go), bit-exact.
//
// Exactness rules: the reference accumulates float64(p)*w in ascending tap order
// with separate multiply and add (Go on amd64 does not fuse), rounds half away from zero into a
// uint8 intermediate after the horizontal pass, and again after the vertical pass.  Every
// arithmetic step that can influence a rounding is issued as __dmul_rn / __dadd_rn (never
// contracted), in the reference's order.
//
// Fast path for GaussianBlur (DESIGN.md "K5"): the 13-tap (sigma=2) sums are first evaluated with
// FP32 FMAs.  Inputs are exact integers <= 255 and the weights are a convex combination, so the FP32
// value is within eps = (taps+2)*255*2^-23 of the FP64 sequence; only when it lies within eps of a
// rounding boundary (k + 0.5) is the exact FP64 sequence evaluated, so the uint8 result is always
// the reference's.  The uint8 intermediate between the passes is kept (effects.go:186-188).
#include "common.cuh"

#include <stdlib.h>

namespace fb {

namespace {

struct BlurParams {
    const uint8_t *src;      // image the taps read
    const uint8_t *alpha;    // image alpha is copied from (always the original source, effects.go:189,215)
    uint8_t *dst;
    long long srcImgStride, alphaImgStride, dstImgStride;
    int srcRowStride, alphaRowStride, dstRowStride;
    int w, h, radius;
    const double *kernel;    // 2r+1 FP64 weights (caller-built)
    const float *kernel32;   // the same rounded to FP32
    float w32[17];           // ... and by value for the fast kernels (radius <= 8): constant-bank operands, no loads
    float eps;
    int exactOnly;
};

__device__ __forceinline__ uint32_t round_fast_or_flag(float v, float eps, bool &ambiguous) {
    float f = v - floorf(v);
    ambiguous = fabsf(f - 0.5f) <= eps;
    float r = floorf(v + 0.5f);
    r = fminf(fmaxf(r, 0.f), 255.f);
    return (uint32_t)r;
}

// One output pixel of a 1-D pass.  VERTICAL=false: effects.go:169-191; true: effects.go:195-217.
template <bool VERTICAL>
__global__ void __launch_bounds__(256) blur_pass_kernel(const BlurParams p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, img = blockIdx.z;
    if (x >= p.w) return;
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;
    const int n = VERTICAL ? p.h : p.w;
    const int pos = VERTICAL ? y : x;
    const int taps = 2 * p.radius + 1;
    auto tap_px = [&](int k) -> uint32_t {
        int q = pos + k - p.radius;
        q = q < 0 ? 0 : (q >= n ? n - 1 : q);  // clamp to edge (effects.go:173-178)
        return VERTICAL ? ld_nc_u32(s + (long long)q * p.srcRowStride + (long long)x * 4)
                        : ld_nc_u32(s + (long long)y * p.srcRowStride + (long long)q * 4);
    };
    uint32_t out[3];
    bool need = p.exactOnly != 0;
    if (!need) {
        float r = 0.f, g = 0.f, b = 0.f;
        for (int k = 0; k < taps; k++) {
            uint32_t v = tap_px(k);
            float wt = __ldg(p.kernel32 + k);
            r = fmaf((float)(v & 0xFF), wt, r);
            g = fmaf((float)((v >> 8) & 0xFF), wt, g);
            b = fmaf((float)((v >> 16) & 0xFF), wt, b);
        }
        bool a0, a1, a2;
        out[0] = round_fast_or_flag(r, p.eps, a0);
        out[1] = round_fast_or_flag(g, p.eps, a1);
        out[2] = round_fast_or_flag(b, p.eps, a2);
        need = a0 | a1 | a2;
    }
    if (need) {  // exact FP64 sequence of the reference
        double r = 0.0, g = 0.0, b = 0.0;
        for (int k = 0; k < taps; k++) {
            uint32_t v = tap_px(k);
            double wt = __ldg(p.kernel + k);
            r = __dadd_rn(r, __dmul_rn((double)(v & 0xFF), wt));
            g = __dadd_rn(g, __dmul_rn((double)((v >> 8) & 0xFF), wt));
            b = __dadd_rn(b, __dmul_rn((double)((v >> 16) & 0xFF), wt));
        }
        out[0] = clampf_dev(r);
        out[1] = clampf_dev(g);
        out[2] = clampf_dev(b);
    }
    uint32_t a = ld_nc_u32(p.alpha + (long long)img * p.alphaImgStride + (long long)y * p.alphaRowStride +
                           (long long)x * 4) & 0xFF000000u;
    *reinterpret_cast<uint32_t *>(p.dst + (long long)img * p.dstImgStride + (long long)y * p.dstRowStride +
                                  (long long)x * 4) = out[0] | (out[1] << 8) | (out[2] << 16) | a;
}

// ------------------------------------------------------------------------------------------------
// Register-tiled fast passes for radius <= 8 (sigma <= 2.66; config 3 uses sigma = 2 → radius 6).
// Each thread produces 16 consecutive outputs along the filter axis, so every input byte is converted
// to float once (PRMT into the mantissa of 2^23, one FADD — I2F runs at 1/8 of the FMA rate) and feeds
// up to 2R+1 accumulators.  All tap indices are compile-time, accumulators never leave registers.
// Outputs whose FP32 value lies within eps of a rounding boundary are recomputed by the same thread in
// the reference's exact FP64 order (~0.1 % of channel values), so bytes stay identical.
// ------------------------------------------------------------------------------------------------
constexpr int kTile = 16;   // outputs per thread
constexpr int kChunkB = 80; // bytes per 16-px chunk in the staging buffer (64 + 16 pad: conflict-free LDS.128)

__device__ __forceinline__ float byte_to_float(uint32_t px, int k) {
    // [byte k, 0x00, 0x00, 0x4B] = bits of 2^23 + byte
    uint32_t m = __byte_perm(px, 0x4B000000u, 0x7540u | (uint32_t)k);
    return __uint_as_float(m) - 8388608.0f;
}

// Round + pack the kTile outputs of a thread.  The accumulators are (even output, odd output) pairs per channel, so
// the magic-number rounding runs as packed FADD2/FFMA2 (t = v + 1.5*2^23, r = t - 1.5*2^23, d = v - r); the low byte of
// t's bit pattern IS the rounded value (v is within [-0.25, 255.25]), so three PRMTs assemble R|G|B|alpha without
// masks or shifts.  Returns the mask of outputs with a channel within eps of a tie (lim = 0.5 - eps).
__device__ __forceinline__ uint32_t pack_rgba_low_bytes(float tr, float tg, float tb, uint32_t alphaWord) {
    const uint32_t x = __byte_perm(__float_as_uint(tr), __float_as_uint(tg), 0x0040);   // [r, g, ., .]
    const uint32_t z = __byte_perm(__float_as_uint(tb), alphaWord, 0x7000);              // [., ., b, alpha]
    return __byte_perm(x, z, 0x7610);
}

template <int TILE, typename AlphaFn>
__device__ __forceinline__ uint32_t blur_round_pack(const float2 (&accRG)[TILE], const float2 (&accB)[TILE / 2], float lim,
                                                    uint32_t (&out)[TILE], AlphaFn alphaWord) {
    const float2 magic = make_float2(12582912.0f, 12582912.0f), nmagic = make_float2(-12582912.0f, -12582912.0f);
    const float2 neg1 = make_float2(-1.0f, -1.0f);
    uint32_t ambMask = 0;
#pragma unroll
    for (int m = 0; m < TILE / 2; m++) {
        float2 t[3], d[3];   // (R,G) of output 2m, (R,G) of output 2m+1, (B, B) of both
        const float2 v[3] = {accRG[2 * m], accRG[2 * m + 1], accB[m]};
#pragma unroll
        for (int c = 0; c < 3; c++) {
            t[c] = __fadd2_rn(v[c], magic);
            const float2 r = __fadd2_rn(t[c], nmagic);
            d[c] = __ffma2_rn(r, neg1, v[c]);   // v - r, exact product
        }
        const float mx = fmaxf(fmaxf(fabsf(d[0].x), fabsf(d[0].y)), fabsf(d[2].x));
        const float my = fmaxf(fmaxf(fabsf(d[1].x), fabsf(d[1].y)), fabsf(d[2].y));
        if (mx >= lim) ambMask |= 1u << (2 * m);
        if (my >= lim) ambMask |= 1u << (2 * m + 1);
        out[2 * m] = pack_rgba_low_bytes(t[0].x, t[0].y, t[2].x, alphaWord(2 * m));
        out[2 * m + 1] = pack_rgba_low_bytes(t[1].x, t[1].y, t[2].y, alphaWord(2 * m + 1));
    }
    return ambMask;
}

// ---- deferred exact path ---------------------------------------------------------------------------------
// About 0.16 % of the outputs are ambiguous, but that is ~1 per warp per 16-output chunk: recomputing them on the
// spot kept one or two lanes busy for ~200 FP64 instructions while the other 30 waited (and the I2F conversions
// queued on the XU pipe: 23 % busy in the round-1c capture, profiles/r1b_ncu_summaries_all_kernels.txt).  Instead each warp queues (x, y) of its ambiguous outputs in
// shared memory and drains the queue 32 at a time, one output per lane, re-reading the taps through L1/L2.
constexpr int kAmbQ = 32 * kTile + 32;   // a chunk can add at most 32*kTile entries to fewer than 32 leftovers

__device__ __forceinline__ double byte_d(uint32_t b) {
    return __hiloint2double(0x43300000, (int)b) - 4503599627370496.0;  // (2^52 + b) - 2^52: exact, no I2F
}

// Exact FP64 sequence (effects.go:172-188 / 198-214) for output (x, y) of a pass; taps are read from `img`
// (row stride `rs`) with clamp-to-edge along the filter axis.  Returns R|G|B; the caller adds alpha.
template <bool VERTICAL>
__device__ __forceinline__ uint32_t blur_exact_at(const uint8_t *img, int rs, int w, int h, int x, int y, int radius,
                                                  const double *kernel, uint32_t &centre) {
    double r = 0.0, g = 0.0, b = 0.0;
    const int n = VERTICAL ? h : w, pos = VERTICAL ? y : x;
    for (int k = 0; k <= 2 * radius; k++) {
        int qd = pos + k - radius;
        qd = qd < 0 ? 0 : (qd >= n ? n - 1 : qd);
        const uint32_t v = VERTICAL ? ld_nc_u32(img + (long long)qd * rs + (long long)x * 4)
                                    : ld_nc_u32(img + (long long)y * rs + (long long)qd * 4);
        if (k == radius) centre = v;
        const double wt = __ldg(kernel + k);
        r = __dadd_rn(r, __dmul_rn(byte_d(v & 0xFF), wt));
        g = __dadd_rn(g, __dmul_rn(byte_d((v >> 8) & 0xFF), wt));
        b = __dadd_rn(b, __dmul_rn(byte_d((v >> 16) & 0xFF), wt));
    }
    return clampf_dev(r) | (clampf_dev(g) << 8) | (clampf_dev(b) << 16);
}

// Queue this lane's flagged outputs: output j of the mask is pixel (x0 + j*dx, y0 + j*dy).
__device__ __forceinline__ void amb_push(uint32_t mask, int x0, int y0, int dx, int dy, uint32_t *q, int *cnt) {
    while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        q[atomicAdd(cnt, 1)] = ((uint32_t)(y0 + j * dy) << 16) | (uint32_t)(x0 + j * dx);
    }
}

// Drain the warp's queue in groups of 32 (all = true: until empty).  Warp-uniform; callers __syncwarp() before.
template <bool VERTICAL>
__device__ __forceinline__ void amb_drain(bool all, int lane, uint32_t *q, int *cnt, const uint8_t *taps, int tapsRs,
                                          uint8_t *dst, int dstRs, int w, int h, int radius, const double *kernel) {
    // lane 0's view, broadcast: the decision must be warp-uniform even if a fast lane has already pushed entries of
    // the next chunk (the shuffle is also the point no lane passes before all have pushed this chunk's entries)
    int n = __shfl_sync(0xffffffffu, *cnt, 0);
    if (n < (all ? 1 : 32)) return;
    while (n >= (all ? 1 : 32)) {
        const int take = min(n, 32);
        n -= take;
        if (lane < take) {
            const uint32_t code = q[n + lane];
            const int x = (int)(code & 0xFFFFu), y = (int)(code >> 16);
            uint32_t centre;
            const uint32_t e = blur_exact_at<VERTICAL>(taps, tapsRs, w, h, x, y, radius, kernel, centre);
            *reinterpret_cast<uint32_t *>(dst + (long long)y * dstRs + (long long)x * 4) = e | (centre & 0xFF000000u);
        }
    }
    __syncwarp();
    if (lane == 0) *cnt = n;
    __syncwarp();
}

// Accumulators: (R, G) of one output share an FFMA2 whose weight is a broadcast scalar; B of outputs (2m, 2m+1)
// sits in the two halves of accB[m] and takes scalar FFMAs.  [The first version paired two OUTPUTS of one channel
// per FFMA2, which needs the weight pairs (w[k], w[k-1]) as aligned register pairs: ptxas rebuilt them with one
// IMAD.MOV per FFMA2 — 28 % of the executed instructions, on the same pipe as the FMAs (profiles/r1d_ncu_summaries_blur_lanczos.txt).]
#ifndef FB_BLUR_I2F
#define FB_BLUR_I2F 2   // 2: every channel through I2F.U8 (XU pipe, otherwise idle): 0.636 ms per 16 4K images against 0.658 with PRMT + FADD; 1: B only, 0.649
#endif
#ifndef FB_BLUR_BA
#define FB_BLUR_BA 0   // 1: B rides an FFMA2 together with the (discarded) alpha lane instead of a scalar FFMA
#endif
template <int R, int NIN, int OFF, int TILE>
__device__ __forceinline__ void blur_taps_fp32(const uint32_t (&raw)[NIN], const float (&w32)[17],
                                               float2 (&accRG)[TILE], float2 (&accB)[TILE / 2]) {
    float wt[2 * R + 1];
#pragma unroll
    for (int k = 0; k <= 2 * R; k++) wt[k] = w32[k];   // kernel parameters: uniform / constant-bank operands
#pragma unroll
    for (int j = 0; j < TILE; j++) accRG[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int m = 0; m < TILE / 2; m++) accB[m] = make_float2(0.f, 0.f);
    const float2 nmagic = make_float2(-8388608.0f, -8388608.0f);
#if FB_BLUR_BA
    float2 accBA[TILE];
#pragma unroll
    for (int j = 0; j < TILE; j++) accBA[j] = make_float2(0.f, 0.f);
#endif
#pragma unroll
    for (int i = OFF - R; i < OFF + TILE + R; i++) {
        // [byte k, 0, 0, 0x4B] = bits of 2^23 + byte; one FADD2 converts R and G, one FADD converts B
#if FB_BLUR_I2F >= 2   // experiment: every channel through I2F.U8 on the XU pipe (16 lanes/clk/SM) instead of PRMT + FADD
        const float2 rg = make_float2((float)(raw[i] & 0xFFu), (float)((raw[i] >> 8) & 0xFFu));
#else
        const float2 rg = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(raw[i], 0x4B000000u, 0x7540u)),
                                                 __uint_as_float(__byte_perm(raw[i], 0x4B000000u, 0x7541u))), nmagic);
#endif
#if FB_BLUR_BA
        const float2 ba = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(raw[i], 0x4B000000u, 0x7542u)),
                                                 __uint_as_float(__byte_perm(raw[i], 0x4B000000u, 0x7543u))), nmagic);
#elif FB_BLUR_I2F >= 1   // experiment: B through I2F.U8 (XU pipe), leaving the FMA pipe one FADD per input less
        const float bl = (float)((raw[i] >> 16) & 0xFFu);
#else
        const float bl = byte_to_float(raw[i], 2);
#endif
#pragma unroll
        for (int j = 0; j < TILE; j++) {
            const int k = i - OFF - j + R;  // tap of input i for output j
            if (k >= 0 && k <= 2 * R) {
                accRG[j] = __ffma2_rn(rg, make_float2(wt[k], wt[k]), accRG[j]);
#if FB_BLUR_BA
                accBA[j] = __ffma2_rn(ba, make_float2(wt[k], wt[k]), accBA[j]);
#else
                if (j & 1) accB[j / 2].y = fmaf(bl, wt[k], accB[j / 2].y);
                else accB[j / 2].x = fmaf(bl, wt[k], accB[j / 2].x);
#endif
            }
        }
    }
#if FB_BLUR_BA
#pragma unroll
    for (int m = 0; m < TILE / 2; m++) accB[m] = make_float2(accBA[2 * m].x, accBA[2 * m + 1].x);
#endif
}

#ifndef FB_BLUR_HROWS
#define FB_BLUR_HROWS 8
#endif
constexpr int kHRows = FB_BLUR_HROWS;  // rows one warp walks in the horizontal pass (double-buffered staging)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Horizontal pass: a warp owns 512 output pixels of a row (16 per lane) and walks kHRows consecutive rows; row
// y+1 is staged into the warp's second shared-memory buffer with cp.async while row y is evaluated, so the
// global-load latency is paid once per warp instead of once per row.  No block-wide barrier.
// WPB = warps per block.  The warps never synchronise with each other, so one-warp blocks only change the granularity
// at which the SM takes on new work (no waiting for the slowest of four warps before the next block starts).
template <int R, int WPB>
__global__ void __launch_bounds__(32 * WPB, 16 / WPB) blur_h_fast_kernel(const BlurParams p) {
    __shared__ __align__(16) uint8_t stage[WPB][2][34 * kChunkB];
    __shared__ uint32_t ambQ[WPB][kAmbQ];
    __shared__ int ambN[WPB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) ambN[warp] = 0;
    __syncwarp();
    const int yBeg = (blockIdx.y * WPB + warp) * kHRows, img = blockIdx.z;
    const int yEnd = min(yBeg + kHRows, p.h);
    if (yBeg >= p.h) return;  // warp-uniform; no block barrier below
    const int xs = blockIdx.x * (32 * kTile);
    const uint8_t *simg = p.src + (long long)img * p.srcImgStride;
    const bool vecOK = ((((uintptr_t)p.src | (uintptr_t)p.srcImgStride | (uintptr_t)p.srcRowStride) & 15) == 0);
    const bool dvec = ((((uintptr_t)p.dst | (uintptr_t)p.dstImgStride | (uintptr_t)p.dstRowStride) & 15) == 0);
    // stage chunks -1..32 (34 chunks of 16 px) of row y with clamp-to-edge (effects.go:173-178)
    // Block-uniform: the whole staged span lies inside the row, so every quad is one aligned 16-byte cp.async
    // (kept separate from the clamped path — if-converted together they cost ~250 address/clamp instructions
    // per warp-row, 18 % of the kernel).
    const bool interiorSpan = vecOK && xs >= kTile && xs + 33 * kTile <= p.w;
    auto stage_row = [&](int y, uint8_t *st) {
        const uint8_t *srow = simg + (long long)y * p.srcRowStride;
        if (interiorSpan) {
            const uint8_t *g0 = srow + (long long)(xs - kTile) * 4 + lane * 16;   // quad v = lane + 32k ↔ 16 bytes at 16*v
            uint8_t *s0 = st + (lane >> 2) * kChunkB + (lane & 3) * 16;
#pragma unroll
            for (int k = 0; k < (34 * 4 + 31) / 32; k++)
                if (k < 4 || lane < 34 * 4 - 128) cp_async16(s0 + k * 8 * kChunkB, g0 + k * 512);
        } else {
            // Edge blocks (the first and the last of a row: 2 of 8 at 3840 px).  Quads that lie inside the row still go
            // through cp.async; only the few that cross an end are loaded by hand with clamp-to-edge.  Unrolled, so the
            // hand loads of different quads are independent.  [A rolled load -> store loop here left every edge warp
            // waiting on four dependent DRAM round trips per row: 23 % of the kernel's stall samples, profiles/r2s2.]
#pragma unroll
            for (int k = 0; k < (34 * 4 + 31) / 32; k++) {
                const int v = lane + 32 * k;
                if (k < 4 || v < 34 * 4) {
                    const int chunk = v >> 2, quad = v & 3;
                    const int px0 = xs + (chunk - 1) * kTile + quad * 4;
                    uint8_t *dstp = st + chunk * kChunkB + quad * 16;
                    if (vecOK && px0 >= 0 && px0 + 4 <= p.w) {
                        cp_async16(dstp, srow + (long long)px0 * 4);
                    } else {
                        uint32_t t[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            int sx = min(max(px0 + i, 0), p.w - 1);
                            t[i] = ld_nc_u32(srow + (long long)sx * 4);
                        }
                        *reinterpret_cast<uint4 *>(dstp) = make_uint4(t[0], t[1], t[2], t[3]);
                    }
                }
            }
        }
        cp_async_commit_group();
    };
    stage_row(yBeg, stage[warp][0]);
    const int x0 = xs + lane * kTile;
    const float lim = 0.5f - p.eps;
#pragma unroll 1
    for (int y = yBeg; y < yEnd; y++) {
        uint8_t *st = stage[warp][(y - yBeg) & 1];
        if (y + 1 < yEnd) {
            stage_row(y + 1, stage[warp][(y + 1 - yBeg) & 1]);
            cp_async_wait_group<1>();
        } else {
            cp_async_wait_group<0>();
        }
        __syncwarp();
        if (x0 < p.w) {
            // window: px x0-8 .. x0+23  =  stage chunks lane (second half), lane+1 (all), lane+2 (first half)
            uint32_t raw[32];
#pragma unroll
            for (int v = 0; v < 8; v++) {
                const int i0 = v * 4 + 8;  // px index relative to the start of stage chunk `lane`
                uint4 q = *reinterpret_cast<const uint4 *>(st + (lane + i0 / kTile) * kChunkB + (i0 % kTile) * 4);
                raw[v * 4 + 0] = q.x; raw[v * 4 + 1] = q.y; raw[v * 4 + 2] = q.z; raw[v * 4 + 3] = q.w;
            }
            float2 accRG[kTile], accB[kTile / 2];
            blur_taps_fp32<R, 32, 8, kTile>(raw, p.w32, accRG, accB);
            uint32_t out[kTile];
            uint32_t ambMask = blur_round_pack<kTile>(accRG, accB, lim, out, [&](int j) { return raw[8 + j]; });  // alpha from the source (effects.go:189)
            uint8_t *drow = p.dst + (long long)img * p.dstImgStride + (long long)y * p.dstRowStride + (long long)x0 * 4;
            if (dvec && x0 + kTile <= p.w) {
#pragma unroll
                for (int v = 0; v < 4; v++)
                    *reinterpret_cast<uint4 *>(drow + v * 16) = make_uint4(out[v * 4], out[v * 4 + 1], out[v * 4 + 2], out[v * 4 + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < kTile; j++)
                    if (x0 + j < p.w) *reinterpret_cast<uint32_t *>(drow + j * 4) = out[j];
            }
            if (x0 + kTile > p.w) ambMask &= (1u << (p.w - x0)) - 1u;
            amb_push(ambMask, x0, y, 1, 0, ambQ[warp], &ambN[warp]);  // exact FP64 path deferred (see amb_drain)
        }
        __syncwarp();  // every lane is done with `st` before it is restaged two rows later; queue pushes are visible
        amb_drain<false>(false, lane, ambQ[warp], &ambN[warp], simg, p.srcRowStride, p.dst + (long long)img * p.dstImgStride,
                         p.dstRowStride, p.w, p.h, R, p.kernel);
    }
    amb_drain<false>(true, lane, ambQ[warp], &ambN[warp], simg, p.srcRowStride, p.dst + (long long)img * p.dstImgStride,
                     p.dstRowStride, p.w, p.h, R, p.kernel);
}

// Vertical pass: a thread owns one column and walks down a segment of kVSeg rows in chunks of kTile
// outputs.  Its window of kTile + 2R packed pixels slides in registers (2R rows are carried over to the next
// chunk) and the kTile new rows of the NEXT chunk are loaded before the current chunk's taps are evaluated, so
// the loads are hidden behind ~300 FFMA2s and every tmp row is read 1 + 2R/kVSeg times instead of
// 1 + 2R/kTile.  Alpha: the horizontal pass has already copied the source alpha into tmp (effects.go:189), so the
// centre tap carries exactly the byte effects.go:215 copies — the original image is not read again.
#ifndef FB_BLUR_VSEG
#define FB_BLUR_VSEG 240
#endif
constexpr int kVSeg = FB_BLUR_VSEG;  // rows per thread segment (2160 = 9 * 240; halo 12/240)

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// The kTile new rows of the NEXT chunk travel through a per-warp shared-memory buffer filled by cp.async (two buffers,
// alternating), not through registers: ptxas put the 16 prefetch LDGs of the register version on the same scoreboard
// as the release of their address registers, so the first instruction of the tap block that reused one of those
// registers waited for ALL the loads — the prefetch was fully exposed (23.6 % of the stall samples on one PRMT,
// profiles/r2s2_blur).  cp.async has no destination registers; its completion is an explicit wait_group at the end of
// the chunk.  Whole-warp-inside-the-row segments copy 16 bytes per lane (8 lanes per row, 4 rows per instruction);
// otherwise every lane copies its own (clamped) column 4 bytes at a time.
#ifndef FB_BLUR_VTILE
#define FB_BLUR_VTILE 16
#endif
#ifndef FB_BLUR_VMINB
#define FB_BLUR_VMINB 16
#endif
constexpr int kTileV = FB_BLUR_VTILE;   // outputs per thread and chunk in the vertical pass
template <int R, int WPB>
__global__ void __launch_bounds__(32 * WPB, FB_BLUR_VMINB / WPB) blur_v_fast_kernel(const BlurParams p) {
    constexpr int NIN = kTileV + 2 * R;
    __shared__ uint32_t ambQ[WPB][kAmbQ];
    __shared__ int ambN[WPB];
    __shared__ __align__(16) uint32_t nxtBuf[WPB][2][kTileV][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) ambN[warp] = 0;
    __syncwarp();
    const int xw = blockIdx.x * (32 * WPB) + warp * 32;   // first column of this warp
    const int x = xw + lane, img = blockIdx.z;
    const int ys = blockIdx.y * kVSeg;
    const int yEnd = min(ys + kVSeg, p.h);
    const bool active = x < p.w;   // inactive lanes still take part in the warp's queue drains
    const int xc = active ? x : p.w - 1;
    const uint8_t *simg = p.src + (long long)img * p.srcImgStride;
    uint8_t *dimg = p.dst + (long long)img * p.dstImgStride;
    const uint8_t *scol = simg + (long long)xc * 4;
    uint8_t *dcol = dimg + (long long)xc * 4;
    const float lim = 0.5f - p.eps;
    const bool wide = xw + 32 <= p.w && ((((uintptr_t)p.src | (uintptr_t)p.srcImgStride | (uintptr_t)p.srcRowStride) & 15) == 0);
    auto ld_row = [&](int y) -> uint32_t {
        const int sy = min(max(y, 0), p.h - 1);  // clamp to edge (effects.go:199-204)
        return __ldg(reinterpret_cast<const uint32_t *>(scol + (long long)sy * p.srcRowStride));
    };
    // rows yn .. yn + kTileV - 1 (clamped to the last row) of this warp's 32 columns into buf[kTileV][32]
    auto stage_rows = [&](int yn, uint32_t (*buf)[32]) {
        if (wide) {
            const uint8_t *g0 = simg + (long long)xw * 4 + (lane & 7) * 16;
#pragma unroll
            for (int k = 0; k < kTileV / 4; k++) {
                const int r = (lane >> 3) + 4 * k;
                cp_async16(&buf[r][(lane & 7) * 4], g0 + (long long)min(yn + r, p.h - 1) * p.srcRowStride);
            }
        } else {
#pragma unroll
            for (int r = 0; r < kTileV; r++) cp_async4(&buf[r][lane], scol + (long long)min(yn + r, p.h - 1) * p.srcRowStride);
        }
        cp_async_commit_group();
    };
    uint32_t raw[NIN];
#pragma unroll
    for (int i = 0; i < NIN; i++) raw[i] = ld_row(ys - R + i);
    int par = 0;
#pragma unroll 1
    for (int y0 = ys; y0 < yEnd; y0 += kTileV, par ^= 1) {
        const bool more = y0 + kTileV < yEnd;
        if (more) stage_rows(y0 + kTileV + R, nxtBuf[warp][par]);   // first new row of the next chunk
        float2 accRG[kTileV], accB[kTileV / 2];
        blur_taps_fp32<R, NIN, R, kTileV>(raw, p.w32, accRG, accB);
        uint32_t out[kTileV];
        uint32_t ambMask = blur_round_pack<kTileV>(accRG, accB, lim, out, [&](int j) { return raw[R + j]; });  // alpha rides in tmp (effects.go:189,215)
        {
            uint8_t *dp = dcol + (long long)y0 * p.dstRowStride;
            if (!active) {
                ambMask = 0;
            } else if (y0 + kTileV <= yEnd) {
#pragma unroll
                for (int j = 0; j < kTileV; j++, dp += p.dstRowStride) *reinterpret_cast<uint32_t *>(dp) = out[j];
            } else {
#pragma unroll
                for (int j = 0; j < kTileV; j++, dp += p.dstRowStride)
                    if (y0 + j < yEnd) *reinterpret_cast<uint32_t *>(dp) = out[j];
                ambMask &= (1u << (yEnd - y0)) - 1u;
            }
        }
        if (active) amb_push(ambMask, x, y0, 0, 1, ambQ[warp], &ambN[warp]);  // exact FP64 path deferred (see amb_drain)
        if (more) cp_async_wait_group<0>();
        __syncwarp();   // queue pushes and (wide) the other lanes' copies are visible; the buffer of the previous chunk is free
        amb_drain<true>(false, lane, ambQ[warp], &ambN[warp], simg, p.srcRowStride, dimg, p.dstRowStride, p.w, p.h, R, p.kernel);
        if (more) {
#pragma unroll
            for (int i = 0; i < 2 * R; i++) raw[i] = raw[i + kTileV];
#pragma unroll
            for (int i = 0; i < kTileV; i++) raw[2 * R + i] = nxtBuf[warp][par][i][lane];
        }
    }
    __syncwarp();
    amb_drain<true>(true, lane, ambQ[warp], &ambN[warp], simg, p.srcRowStride, dimg, p.dstRowStride, p.w, p.h, R, p.kernel);
}

template <int R>
static void launch_blur_fast(cudaStream_t s, BlurParams p, int n, bool vertical) {
    static const bool wpb4 = [] { const char *e = getenv("FB_BLUR_WPB"); return e && e[0] == '4'; }();   // round-1 block shape
    if (!vertical) {
        if (wpb4) {
            dim3 grid((p.w + 32 * kTile - 1) / (32 * kTile), (p.h + 4 * kHRows - 1) / (4 * kHRows), n);
            blur_h_fast_kernel<R, 4><<<grid, 128, 0, s>>>(p);
        } else {
            dim3 grid((p.w + 32 * kTile - 1) / (32 * kTile), (p.h + kHRows - 1) / kHRows, n);
            blur_h_fast_kernel<R, 1><<<grid, 32, 0, s>>>(p);
        }
    } else {
        if (wpb4) {
            dim3 grid((p.w + 127) / 128, (p.h + kVSeg - 1) / kVSeg, n);
            blur_v_fast_kernel<R, 4><<<grid, 128, 0, s>>>(p);
        } else {
            dim3 grid((p.w + 31) / 32, (p.h + kVSeg - 1) / kVSeg, n);
            blur_v_fast_kernel<R, 1><<<grid, 32, 0, s>>>(p);
        }
    }
}

static bool launch_blur_fast_any(cudaStream_t s, const BlurParams &p, int n, bool vertical) {
    switch (p.radius) {
        case 1: launch_blur_fast<1>(s, p, n, vertical); return true;
        case 2: launch_blur_fast<2>(s, p, n, vertical); return true;
        case 3: launch_blur_fast<3>(s, p, n, vertical); return true;
        case 4: launch_blur_fast<4>(s, p, n, vertical); return true;
        case 5: launch_blur_fast<5>(s, p, n, vertical); return true;
        case 6: launch_blur_fast<6>(s, p, n, vertical); return true;
        case 7: launch_blur_fast<7>(s, p, n, vertical); return true;
        case 8: launch_blur_fast<8>(s, p, n, vertical); return true;
        default: return false;
    }
}

struct FxParams {
    const uint8_t *src;
    uint8_t *dst;
    long long srcImgStride, dstImgStride;
    int srcRowStride, dstRowStride;
    int w, h;
    double amount;
    int mode;  // 0 = blur3x3 only, 1 = Sharpen, 2 = AdaptiveSharpen
};

// gaussianBlur3x3 value of one channel triple at an interior pixel: clampF(sum/16) == (sum+8)>>4
// because sum is an integer <= 4080 and /16 is exact (effects.go:124-136).
__device__ __forceinline__ void blur3_rgb(const uint32_t (&n)[9], uint32_t (&out)[3]) {
    const int wt[9] = {1, 2, 1, 2, 4, 2, 1, 2, 1};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) s += ((n[k] >> (8 * c)) & 0xFF) * wt[k];
        out[c] = (s + 8) >> 4;
    }
}

__device__ __forceinline__ double luma64(uint32_t v) {  // effects.go:96
    return __dadd_rn(__dadd_rn(__dmul_rn(0.299, (double)(v & 0xFF)), __dmul_rn(0.587, (double)((v >> 8) & 0xFF))),
                     __dmul_rn(0.114, (double)((v >> 16) & 0xFF)));
}

// effects.go:93-112 with the neighbourhood n[ky*3+kx], expression order preserved.
__device__ __forceinline__ double edge_strength(const uint32_t (&n)[9]) {
    double l00 = luma64(n[0]), l10 = luma64(n[1]), l20 = luma64(n[2]);
    double l01 = luma64(n[3]), l21 = luma64(n[5]);
    double l02 = luma64(n[6]), l12 = luma64(n[7]), l22 = luma64(n[8]);
    double gx = __dadd_rn(-l00, l20);
    gx = __dadd_rn(gx, -__dmul_rn(2.0, l01));
    gx = __dadd_rn(gx, __dmul_rn(2.0, l21));
    gx = __dadd_rn(gx, -l02);
    gx = __dadd_rn(gx, l22);
    double gy = __dadd_rn(-l00, -__dmul_rn(2.0, l10));
    gy = __dadd_rn(gy, -l20);
    gy = __dadd_rn(gy, l02);
    gy = __dadd_rn(gy, __dmul_rn(2.0, l12));
    gy = __dadd_rn(gy, l22);
    double mag = __dsqrt_rn(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)));
    double normalized = __ddiv_rn(mag, 400.0);
    return normalized > 1.0 ? 1.0 : normalized;
}

// Fused 3x3 blur + unsharp: the blurred image is never materialised (8 B/px of HBM traffic).
__global__ void __launch_bounds__(256) sharpen_kernel(const FxParams p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, img = blockIdx.z;
    if (x >= p.w) return;
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;
    const uint32_t centre = ld_nc_u32(s + (long long)y * p.srcRowStride + (long long)x * 4);
    uint32_t result = centre;
    const bool interior = x >= 1 && x < p.w - 1 && y >= 1 && y < p.h - 1;
    if (interior) {
        uint32_t n[9];
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
            for (int kx = 0; kx < 3; kx++)
                n[ky * 3 + kx] = ld_nc_u32(s + (long long)(y + ky - 1) * p.srcRowStride + (long long)(x + kx - 1) * 4);
        uint32_t bl[3];
        blur3_rgb(n, bl);
        if (p.mode == 0) {
            result = bl[0] | (bl[1] << 8) | (bl[2] << 16) | (centre & 0xFF000000u);
        } else {
            double amount = p.amount;
            if (p.mode == 2) amount = __dmul_rn(p.amount, edge_strength(n));  // effects.go:73-74
            uint32_t o[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                int orig = (centre >> (8 * c)) & 0xFF;
                double val = __dadd_rn((double)orig, __dmul_rn(amount, (double)(orig - (int)bl[c])));  // :37 / :82
                o[c] = clampf_dev(val);
            }
            result = o[0] | (o[1] << 8) | (o[2] << 16) | (centre & 0xFF000000u);
        }
    }
    // Border pixels: blur == src there, so Sharpen yields clampF(orig + amount*0) == orig
    // (effects.go:28-42); AdaptiveSharpen and blur3x3 copy the source (effects.go:68,120).
    *reinterpret_cast<uint32_t *>(p.dst + (long long)img * p.dstImgStride + (long long)y * p.dstRowStride +
                                  (long long)x * 4) = result;
}

// ------------------------------------------------------------------------------------------------
// Tiled Sharpen / AdaptiveSharpen / blur3x3: each thread owns 4 adjacent pixels (one 128-bit load per
// row) and walks kFxRows rows down, keeping the horizontal 1-2-1 sums of the previous two rows in
// registers, so every source pixel is loaded ~1.5 times instead of 9.  The 3x3 sums are SIMD-in-register
// on 16-bit lanes (R|B and G|A words).  Finish per channel:
//   INTK >= 0: amount * 2^INTK is an integer A (e.g. 1.75 = 7/4): V = (orig << k) + A*(orig - blur) is the
//              reference's FP64 value times 2^k EXACTLY, so round-half-away is (V + 2^(k-1)) >> k.
//   INTK <  0: FP64 with the reference's multiply/add order; int -> double through the 2^52 mantissa
//              trick and rounding through u = x + 2^52 (no I2F/F2I/FRND: those run at 16/clk/SM).
// ------------------------------------------------------------------------------------------------
constexpr int kFxRows = 8;

__device__ __forceinline__ double small_int_to_double(int v) {  // exact for |v| < 2^31
    return __hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)) - 4503601774854144.0;  // 2^52 + 2^31
}

// clampF (convert.go:149-158) for |x| < 2^31 without conversion instructions.
__device__ __forceinline__ uint32_t clampf_magic(double x) {
    if (!(x >= 0.5)) return 0u;        // (-inf, 0.5) rounds to <= 0
    if (x >= 254.5) return 255u;
    double u = x + 4503599627370496.0;  // RNE to integer in the low mantissa bits
    double t = u - 4503599627370496.0;
    uint32_t r = (uint32_t)__double2loint(u);
    if (x - t == 0.5) r += 1;           // RNE went down on a tie; the reference rounds half away from zero
    return r;
}

struct FxTileParams {
    const uint8_t *src;
    uint8_t *dst;
    long long srcImgStride, dstImgStride;
    int srcRowStride, dstRowStride;
    int w, h;
    double amount;
    int A, half;   // integer path: A = amount * 2^k, half = 2^(k-1)
    int vecOK;
    int fastTiles; // Sharpen: interior tiles take sharpen_tile_fast (0: FB_FX_NOFAST=1, the round-1 path everywhere)
};

// Sharpen with a dyadic amount on a tile that lies strictly inside the image (no clamps, no per-pixel border tests, every
// row an aligned 128-bit load plus the two neighbour pixels at immediate offsets): the same integer arithmetic as the
// INTK >= 0 branch of fx_tile_kernel below, ~35 instead of ~62 instructions per pixel.  [fx_tile_kernel<1, 2> executed 70
// instructions per pixel with the ALU pipe 78 % busy (profiles/r1d_ncu_summaries_msssim_analyze_fx.txt): border tests and
// their branches, window-rotation moves, shift+mask splits and the shift/or/and packing were half of it.]  Splits use one
// PRMT per 16-bit-lane word, the output is assembled with two PRMTs, the three-row window is renamed by full unrolling.
#ifndef FB_FX_PF
#define FB_FX_PF 2
#endif
#ifndef FB_FX_IMAD
#define FB_FX_IMAD 1
#endif
template <int K>
__device__ __forceinline__ void sharpen_tile_fast(const FxTileParams &p, const uint8_t *s, uint8_t *d, int x0, int yb) {
    const uint32_t cst = (uint32_t)((1024 << K) + p.half) * 0x00010001u;
    const uint32_t msk = (uint32_t)(0xFFFF >> K) * 0x00010001u;
    const uint32_t mulO = (uint32_t)(p.A + (1 << K)), mulB = (uint32_t)p.A;
    const uint32_t two = (uint32_t)p.fastTiles + 1u;   // == 2 on this path, from the parameter bank
    const uint32_t sh4 = two << 27;   // 2^28: __umulhi(x, 2^28) == x >> 4 on the FMA pipe
    (void)two; (void)sh4;
    const uint8_t *row = s + (long long)(yb - 1) * p.srcRowStride + (long long)x0 * 4;
    uint8_t *drow = d + (long long)yb * p.dstRowStride + (long long)x0 * 4;
    // horizontal 1-2-1 sums of a row on packed 16-bit lanes (R|B and G|A words) + the row's own pixels split the same way
    auto load_hsum = [&](const uint8_t *r, uint32_t (&hrb)[4], uint32_t (&hga)[4], uint32_t (&orb)[4], uint32_t (&oga)[4], uint32_t (&raw)[4]) {
        const uint4 q = ld_nc_u128(r);   // non-coherent like the neighbour loads: free to move above the previous rows' stores
        const uint32_t px[6] = {ld_nc_u32(r - 4), q.x, q.y, q.z, q.w, ld_nc_u32(r + 16)};
        uint32_t rb[6], ga[6];
#pragma unroll
        for (int i = 0; i < 6; i++) { rb[i] = px[i] & 0x00FF00FFu; ga[i] = __byte_perm(px[i], 0u, 0x4341u); }
#pragma unroll
        for (int i = 0; i < 4; i++) {
#if FB_FX_IMAD   // x + 2y as an IMAD with the 2 from the parameter bank (ptxas cannot turn it back into an ALU-pipe LEA)
            hrb[i] = rb[i + 1] * two + rb[i] + rb[i + 2];
            hga[i] = ga[i + 1] * two + ga[i] + ga[i + 2];
#else
            hrb[i] = rb[i] + 2 * rb[i + 1] + rb[i + 2];
            hga[i] = ga[i] + 2 * ga[i + 1] + ga[i + 2];
#endif
            orb[i] = rb[i + 1]; oga[i] = ga[i + 1]; raw[i] = px[i + 1];
        }
    };
#if FB_FX_PF
    // L2 prefetch of the tile's later rows: 72 registers hold two or three rows of loads in flight, not ten
#pragma unroll
    for (int r = FB_FX_PF; r < kFxRows + 2; r++) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + (long long)r * p.srcRowStride));
#endif
    uint32_t hP_rb[4], hP_ga[4], hC_rb[4], hC_ga[4], hN_rb[4], hN_ga[4];
    uint32_t oC_rb[4], oC_ga[4], rawC[4], oN_rb[4], oN_ga[4], rawN[4];
    load_hsum(row, hP_rb, hP_ga, oN_rb, oN_ga, rawN);               // row yb-1 (its own pixels are not needed)
    load_hsum(row + p.srcRowStride, hC_rb, hC_ga, oC_rb, oC_ga, rawC);   // row yb
    row += 2 * (long long)p.srcRowStride;
#pragma unroll
    for (int r = 0; r < kFxRows; r++, row += p.srcRowStride, drow += p.dstRowStride) {
        load_hsum(row, hN_rb, hN_ga, oN_rb, oN_ga, rawN);           // row yb + r + 1
        uint32_t out[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            // gaussianBlur3x3 (effects.go:124-136): (sum + 8) >> 4 on each 16-bit lane
#if FB_FX_IMAD >= 2   // ... and the >> 4 as IMAD.HI by 2^28
            const uint32_t brb = __umulhi(hC_rb[i] * two + hP_rb[i] + hN_rb[i] + 0x00080008u, sh4) & 0x00FF00FFu;
            const uint32_t bga = __umulhi(hC_ga[i] * two + hP_ga[i] + hN_ga[i] + 0x00080008u, sh4) & 0x00FF00FFu;
#elif FB_FX_IMAD
            const uint32_t brb = ((hC_rb[i] * two + hP_rb[i] + hN_rb[i] + 0x00080008u) >> 4) & 0x00FF00FFu;
            const uint32_t bga = ((hC_ga[i] * two + hP_ga[i] + hN_ga[i] + 0x00080008u) >> 4) & 0x00FF00FFu;
#else
            const uint32_t brb = ((hP_rb[i] + 2 * hC_rb[i] + hN_rb[i] + 0x00080008u) >> 4) & 0x00FF00FFu;
            const uint32_t bga = ((hP_ga[i] + 2 * hC_ga[i] + hN_ga[i] + 0x00080008u) >> 4) & 0x00FF00FFu;
#endif
            // integer unsharp, see fx_tile_kernel: T = orig*(A + 2^k) - blur*A + half + bias; out = relu(min((T >> k) - 1024, 255))
            const uint32_t trb = ((oC_rb[i] * mulO + cst - brb * mulB) >> K) & msk;
            const uint32_t tga = ((oC_ga[i] * mulO + cst - bga * mulB) >> K) & msk;
            const uint32_t rrb = __viaddmin_s16x2_relu(trb, 0xFC00FC00u, 0x00FF00FFu);   // [R, 0, B, 0]
            const uint32_t rga = __viaddmin_s16x2_relu(tga, 0xFC00FC00u, 0x00FF00FFu);   // [G, 0, x, 0]
            out[i] = __byte_perm(__byte_perm(rrb, rga, 0x3240u), rawC[i], 0x7210u);       // [R, G, B, alpha of the source]
        }
        *reinterpret_cast<uint4 *>(drow) = make_uint4(out[0], out[1], out[2], out[3]);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            hP_rb[i] = hC_rb[i]; hP_ga[i] = hC_ga[i]; hC_rb[i] = hN_rb[i]; hC_ga[i] = hN_ga[i];
            oC_rb[i] = oN_rb[i]; oC_ga[i] = oN_ga[i]; rawC[i] = rawN[i];
        }
    }
}

template <int MODE /*0 blur3x3, 1 sharpen, 2 adaptive*/, int INTK>
__global__ void __launch_bounds__(128) fx_tile_kernel(const FxTileParams p) {
    const int x0 = (blockIdx.x * 128 + threadIdx.x) * 4;
    const int yb = blockIdx.y * kFxRows, img = blockIdx.z;
    if (x0 >= p.w) return;
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;
    uint8_t *d = p.dst + (long long)img * p.dstImgStride;
    const bool full = p.vecOK && x0 + 4 <= p.w;
    if (MODE == 1 && INTK >= 0) {
        // warp-uniform: every tile of the warp strictly inside the image, source and destination rows 16-byte aligned
        const bool dal = (((uintptr_t)p.dst | (uintptr_t)p.dstImgStride | (uintptr_t)p.dstRowStride) & 15) == 0;
        const bool inside = full && dal && x0 >= 4 && x0 + 5 <= p.w && yb >= 1 && yb + kFxRows + 1 <= p.h;
        if (__all_sync(__activemask(), inside) && p.fastTiles) {
            sharpen_tile_fast<(INTK >= 0 ? INTK : 0)>(p, s, d, x0, yb);
            return;
        }
    }
    const int xl = max(x0 - 1, 0), xr = min(x0 + 4, p.w - 1);

    // px[0] = left neighbour, px[1..4] = own pixels, px[5] = right neighbour (clamped: only borders see the clamp)
    auto load_row = [&](int y, uint32_t (&px)[6]) {
        const uint8_t *row = s + (long long)min(max(y, 0), p.h - 1) * p.srcRowStride;
        if (full) {
            uint4 q = *reinterpret_cast<const uint4 *>(row + (long long)x0 * 4);
            px[1] = q.x; px[2] = q.y; px[3] = q.z; px[4] = q.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) px[1 + i] = ld_nc_u32(row + (long long)min(x0 + i, p.w - 1) * 4);
        }
        px[0] = ld_nc_u32(row + (long long)xl * 4);
        px[5] = ld_nc_u32(row + (long long)xr * 4);
    };
    // horizontal 1-2-1 sums on packed 16-bit lanes: hrb = R|B, hga = G|A
    auto hsum = [&](const uint32_t (&px)[6], uint32_t (&hrb)[4], uint32_t (&hga)[4]) {
        uint32_t rb[6], ga[6];
#pragma unroll
        for (int i = 0; i < 6; i++) { rb[i] = px[i] & 0x00FF00FFu; ga[i] = (px[i] >> 8) & 0x00FF00FFu; }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            hrb[i] = rb[i] + 2 * rb[i + 1] + rb[i + 2];
            hga[i] = ga[i] + 2 * ga[i + 1] + ga[i + 2];
        }
    };

    // MODE 2 fast path: integer lumas x1000 (299R + 587G + 114B, exact) of the three live rows
    auto lumas = [&](const uint32_t (&px)[6], int (&L)[6]) {
#pragma unroll
        for (int i = 0; i < 6; i++)
            L[i] = (int)__dp2a_lo(299u | (587u << 16), px[i], __dp2a_hi(114u, px[i], 0u));
    };
    const float amountF = (float)p.amount;

    uint32_t pPrev[6], pCur[6], pNext[6];
    uint32_t hPrevRB[4], hPrevGA[4], hCurRB[4], hCurGA[4], hNextRB[4], hNextGA[4];
    int lPrev[6], lCur[6], lNext[6];
    load_row(yb - 1, pPrev);
    load_row(yb, pCur);
    hsum(pPrev, hPrevRB, hPrevGA);
    hsum(pCur, hCurRB, hCurGA);
    if (MODE == 2) { lumas(pPrev, lPrev); lumas(pCur, lCur); }
    // MODE 0/1: fully unrolled so the three-row window is renamed instead of rotated with ~28 MOVs per row
    // (the MODE 2 body is too large to replicate eight times).
#pragma unroll (MODE == 2 ? 1 : kFxRows)
    for (int r = 0; r < kFxRows; r++) {
        const int y = yb + r;
        if (y >= p.h) break;
        load_row(y + 1, pNext);
        hsum(pNext, hNextRB, hNextGA);
        int colsum[6];
        if (MODE == 2) {
            lumas(pNext, lNext);
#pragma unroll
            for (int i = 0; i < 6; i++) colsum[i] = lPrev[i] + 2 * lCur[i] + lNext[i];
        }
        uint32_t out[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int x = x0 + i;
            const uint32_t c = pCur[1 + i];
            uint32_t res = c;
            const bool interior = x >= 1 && x < p.w - 1 && y >= 1 && y < p.h - 1;
            if (interior) {
                // gaussianBlur3x3 (effects.go:124-136): (sum + 8) >> 4 on each 16-bit lane
                const uint32_t brb = ((hPrevRB[i] + 2 * hCurRB[i] + hNextRB[i] + 0x00080008u) >> 4) & 0x00FF00FFu;
                const uint32_t bga = ((hPrevGA[i] + 2 * hCurGA[i] + hNextGA[i] + 0x00080008u) >> 4) & 0x00FF00FFu;
                if (MODE == 1 && INTK >= 0) {
                    constexpr int K = INTK >= 0 ? INTK : 0;      // INTK < 0 never reaches here; keeps the shifts well-formed
                    // Integer unsharp on packed 16-bit lanes (R|B and G|A): per lane
                    //   T = orig*(A + 2^k) - blur*A + half + bias,   bias = 1024 << k  (lanes stay in [0, 32767])
                    //   out = relu(min((T >> k) - 1024, 255))        one DPX instruction (VIADDMNMX.S16x2.RELU)
                    // == min(max(((orig << k) + A*(orig - blur) + half) >> k, 0), 255), the exact form of effects.go:37-38
                    // for dyadic amounts (checked for every (orig, blur) pair and every (A, k) the launcher can pick).
                    const uint32_t cst = (uint32_t)((1024 << K) + p.half) * 0x00010001u;
                    const uint32_t msk = (uint32_t)(0xFFFF >> K) * 0x00010001u;
                    const uint32_t orb = c & 0x00FF00FFu, oga = (c >> 8) & 0x00FF00FFu;
                    const uint32_t trb = ((orb * (uint32_t)(p.A + (1 << K)) + cst - brb * (uint32_t)p.A) >> K) & msk;
                    const uint32_t tga = ((oga * (uint32_t)(p.A + (1 << K)) + cst - bga * (uint32_t)p.A) >> K) & msk;
                    const uint32_t rrb = __viaddmin_s16x2_relu(trb, 0xFC00FC00u, 0x00FF00FFu);   // + (-1024) per lane
                    const uint32_t rga = __viaddmin_s16x2_relu(tga, 0xFC00FC00u, 0x00FF00FFu);
                    out[i] = ((rrb | (rga << 8)) & 0x00FFFFFFu) | (c & 0xFF000000u);
                    continue;
                }
                const int bl[3] = {(int)(brb & 0xFF), (int)(bga & 0xFF), (int)(brb >> 16)};
                if (MODE == 0) {
                    res = (uint32_t)bl[0] | ((uint32_t)bl[1] << 8) | ((uint32_t)bl[2] << 16) | (c & 0xFF000000u);
                } else {
                    bool needExact = true;
                    uint32_t o[3];
                    if (MODE == 2) {
                        // FP32 evaluation with a rigorous error bound; only results within the bound of a rounding
                        // boundary take the exact FP64 sequence below.  GX, GY are the Sobel sums of the integer lumas
                        // (units of 1/1000; |G| <= 1.02e6 < 2^22 so the magic-number conversion is exact; the reference's
                        // FP64 gx, gy differ from G/1000 by ~1e-12).  Relative error of edge <= 2^-22 (g2: 2^-23, IEEE
                        // sqrt: halves it + 2^-24, x fl(1/400000): 2^-23), of la*diff <= 3.5 * 2^-23, |la*diff| <= 765;
                        // the final add rounds once more (<= 1020 * 2^-24).  eps = |t| * 5.3e-7 + 8e-5 has 25 % margin.
                        const int GX = colsum[i + 2] - colsum[i];
                        const int GY = (lNext[i] + 2 * lNext[i + 1] + lNext[i + 2]) - (lPrev[i] + 2 * lPrev[i + 1] + lPrev[i + 2]);
                        const float gxf = __int_as_float(GX + 0x4B400000) - 12582912.0f;
                        const float gyf = __int_as_float(GY + 0x4B400000) - 12582912.0f;
                        const float g2f = fmaf(gxf, gxf, gyf * gyf);
                        const float edgeF = fminf(__fsqrt_rn(g2f) * 2.5e-6f, 1.0f);
                        const float la = amountF * edgeF;
                        bool amb = false;
#pragma unroll
                        for (int ch = 0; ch < 3; ch++) {
                            const int orig = (int)((c >> (8 * ch)) & 0xFF);
                            const int diff = orig - bl[ch];
                            const float of = __int_as_float(orig + 0x4B400000) - 12582912.0f;
                            const float df = __int_as_float(diff + 0x4B400000) - 12582912.0f;
                            const float t = la * df;
                            float v = of + t;
                            const float lim = 0.5f - fmaf(fabsf(t), 5.3e-7f, 8e-5f);
                            v = fminf(fmaxf(v, -1.0f), 256.0f);
                            const float tt = v + 12582912.0f;
                            const float rounded = tt - 12582912.0f;
                            amb |= fabsf(v - rounded) >= lim;
                            const int iv = (int)(__float_as_uint(tt) & 0x7FFFFFu) - 0x400000;
                            o[ch] = (uint32_t)min(max(iv, 0), 255);
                        }
                        needExact = amb;
                    }
                    if (needExact) {
                    double amount = p.amount;
                    if (MODE == 2) {  // localEdgeStrength (effects.go:93-112), expression order preserved
                        const uint32_t n[9] = {pPrev[i], pPrev[i + 1], pPrev[i + 2], pCur[i], c, pCur[i + 2],
                                               pNext[i], pNext[i + 1], pNext[i + 2]};
                        auto lum = [](uint32_t v) {
                            return __dadd_rn(__dadd_rn(__dmul_rn(0.299, small_int_to_double((int)(v & 0xFF))),
                                                       __dmul_rn(0.587, small_int_to_double((int)((v >> 8) & 0xFF)))),
                                             __dmul_rn(0.114, small_int_to_double((int)((v >> 16) & 0xFF))));
                        };
                        const double l00 = lum(n[0]), l10 = lum(n[1]), l20 = lum(n[2]), l01 = lum(n[3]), l21 = lum(n[5]);
                        const double l02 = lum(n[6]), l12 = lum(n[7]), l22 = lum(n[8]);
                        double gx = __dadd_rn(-l00, l20);
                        gx = __dadd_rn(gx, -__dmul_rn(2.0, l01));
                        gx = __dadd_rn(gx, __dmul_rn(2.0, l21));
                        gx = __dadd_rn(gx, -l02);
                        gx = __dadd_rn(gx, l22);
                        double gy = __dadd_rn(-l00, -__dmul_rn(2.0, l10));
                        gy = __dadd_rn(gy, -l20);
                        gy = __dadd_rn(gy, l02);
                        gy = __dadd_rn(gy, __dmul_rn(2.0, l12));
                        gy = __dadd_rn(gy, l22);
                        const double g2 = __dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy));
                        double edge = 1.0;
                        // sqrt and /400 are monotone and correctly rounded: g2 >= 160001 gives mag/400 > 1 → clamped to 1
                        if (g2 < 160001.0) {
                            edge = __ddiv_rn(__dsqrt_rn(g2), 400.0);
                            if (edge > 1.0) edge = 1.0;
                        }
                        amount = __dmul_rn(p.amount, edge);  // effects.go:74
                    }
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        const int orig = (int)((c >> (8 * ch)) & 0xFF);
                        const int diff = orig - bl[ch];
                        if (INTK >= 0) {
                            constexpr int K = INTK >= 0 ? INTK : 0;
                            int V = (orig << K) + p.A * diff + p.half;   // exact (see header comment)
                            V = V < 0 ? 0 : (V >> K);
                            o[ch] = (uint32_t)min(V, 255);
                        } else {
                            double val = __dadd_rn(small_int_to_double(orig), __dmul_rn(amount, small_int_to_double(diff)));
                            o[ch] = clampf_magic(val);  // effects.go:37-38 / 82-83
                        }
                    }
                    }
                    res = o[0] | (o[1] << 8) | (o[2] << 16) | (c & 0xFF000000u);
                }
            }
            out[i] = res;
        }
        uint8_t *drow = d + (long long)y * p.dstRowStride + (long long)x0 * 4;
        if (full && ((((uintptr_t)p.dst | (uintptr_t)p.dstImgStride | (uintptr_t)p.dstRowStride) & 15) == 0)) {
            *reinterpret_cast<uint4 *>(drow) = make_uint4(out[0], out[1], out[2], out[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (x0 + i < p.w) *reinterpret_cast<uint32_t *>(drow + i * 4) = out[i];
        }
#pragma unroll
        for (int i = 0; i < 6; i++) { pPrev[i] = pCur[i]; pCur[i] = pNext[i]; }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 6; i++) { lPrev[i] = lCur[i]; lCur[i] = lNext[i]; }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) { hPrevRB[i] = hCurRB[i]; hPrevGA[i] = hCurGA[i]; hCurRB[i] = hNextRB[i]; hCurGA[i] = hNextGA[i]; }
    }
}

// ------------------------------------------------------------------------------------------------
// AdaptiveSharpen, second version.  fx_tile_kernel<2> (kept as FB_ADAPTIVE_OLD=1) spent 170 instructions per pixel:
// four inlined copies of the FP64 reference sequence for the rare ambiguous pixels plus ~130 for the fast path
// (IEEE sqrt, per-channel bounds, clamps on both sides of the rounding).  Here
//   * ambiguous pixels are only QUEUED per warp ((x, y) in 16+16 bits) and the exact FP64 sequence of
//     effects.go:49-112 runs 32 at a time, one pixel per lane, re-reading the 3x3 neighbourhood through L1/L2;
//   * |grad| = g2 * rsqrt(g2) (MUFU, <= 2 ulp), one bound per pixel, v = fma(la, diff, orig) with a single rounding,
//     clamp to [0, 255] BEFORE the magic-number rounding so the low byte of the bit pattern is the result.
// Error bound (relative, first order): g2 2^-23 -> sqrt halves it; rsqrt 2^-22, product 2^-24  =>  mag 3.6e-7;
// x fl(1/400000): 4.8e-7; x fl(amount): 6.0e-7 on la; |la*diff| <= 765; the fma rounds once (<= 1020 * 2^-24).
// eps = la * 255 * 7.5e-7 + 8e-5 (25 % margin on the first term) is a bound for every channel of the pixel.
// ------------------------------------------------------------------------------------------------
constexpr int kAdQ = 128 + 32;   // a warp adds at most 128 entries per row to fewer than 32 leftovers

__device__ __forceinline__ uint32_t adaptive_exact_at(const uint8_t *img, int rs, int x, int y, double amount0) {
    uint32_t n[9];
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++) n[ky * 3 + kx] = ld_nc_u32(img + (long long)(y - 1 + ky) * rs + (long long)(x - 1 + kx) * 4);
    uint32_t bl[3];
    blur3_rgb(n, bl);
    const double amount = __dmul_rn(amount0, edge_strength(n));   // effects.go:74
    const uint32_t c = n[4];
    uint32_t o[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const int orig = (int)((c >> (8 * ch)) & 0xFF);
        const double val = __dadd_rn(small_int_to_double(orig), __dmul_rn(amount, small_int_to_double(orig - (int)bl[ch])));
        o[ch] = clampf_magic(val);   // effects.go:82-83
    }
    return o[0] | (o[1] << 8) | (o[2] << 16) | (c & 0xFF000000u);
}

// AdaptiveSharpen on a tile strictly inside the image (the counterpart of sharpen_tile_fast): aligned non-coherent row
// loads with an L2 prefetch of the later rows, PRMT splits, no border tests, the horizontal 1-2-1 luma sums shared by the
// Sobel rows above and below.  Same FP32 evaluation, same bound and the same per-warp exact queue as the general body
// below; returns through the caller's queue drain.
#ifndef FB_AD_UNROLL
#define FB_AD_UNROLL 1   // rows of the lean AdaptiveSharpen loop unrolled (the three-row window is renamed instead of moved)
#endif
constexpr int kAdUnroll = FB_AD_UNROLL;
__device__ __forceinline__ uint32_t adaptive_row_fast(const uint32_t (&hP_rb)[4], const uint32_t (&hP_ga)[4], const uint32_t (&hC_rb)[4],
                                                      const uint32_t (&hC_ga)[4], const uint32_t (&hN_rb)[4], const uint32_t (&hN_ga)[4],
                                                      const int (&hlP)[4], const int (&hlN)[4], const int (&colsum)[6],
                                                      const uint32_t (&rawC)[4], float amountF, uint32_t (&out)[4]) {
    const float kMagic = 12582912.0f;
    uint32_t ambMask = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t c = rawC[i];
        const uint32_t brb = ((hP_rb[i] + 2 * hC_rb[i] + hN_rb[i] + 0x00080008u) >> 4) & 0x00FF00FFu;
        const uint32_t bga = ((hP_ga[i] + 2 * hC_ga[i] + hN_ga[i] + 0x00080008u) >> 4) & 0x00FF00FFu;
        const int GX = colsum[i + 2] - colsum[i];
        const int GY = hlN[i] - hlP[i];
        const float gxf = __int_as_float(GX + 0x4B400000) - kMagic;
        const float gyf = __int_as_float(GY + 0x4B400000) - kMagic;
        const float g2 = fmaxf(fmaf(gxf, gxf, gyf * gyf), 1e-30f);
        const float la = amountF * fminf(g2 * rsqrtf(g2) * 2.5e-6f, 1.0f);
        const float lim = 0.5f - fmaf(la, 255.0f * 7.5e-7f, 8e-5f);
        const float oR = __uint_as_float(__byte_perm(c, 0x4B400000u, 0x7650u)) - kMagic;
        const float oG = __uint_as_float(__byte_perm(c, 0x4B400000u, 0x7651u)) - kMagic;
        const float oB = __uint_as_float(__byte_perm(c, 0x4B400000u, 0x7652u)) - kMagic;
        const float bR = __uint_as_float(__byte_perm(brb, 0x4B400000u, 0x7650u)) - kMagic;
        const float bG = __uint_as_float(__byte_perm(bga, 0x4B400000u, 0x7650u)) - kMagic;
        const float bB = __uint_as_float(__byte_perm(brb, 0x4B400000u, 0x7652u)) - kMagic;
        const float o3[3] = {oR, oG, oB}, b3[3] = {bR, bG, bB};
        float t3[3];
        float worst = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            float v = fmaf(la, o3[ch] - b3[ch], o3[ch]);   // orig - blur is an exact small integer
            v = fminf(fmaxf(v, 0.0f), 255.0f);             // ties at -0.5 / 255.5 cannot change the clamped result
            t3[ch] = v + kMagic;
            worst = fmaxf(worst, fabsf(v - (t3[ch] - kMagic)));
        }
        if (worst >= lim) ambMask |= 1u << i;
        out[i] = pack_rgba_low_bytes(t3[0], t3[1], t3[2], c);
    }
    return ambMask;
}

__global__ void __launch_bounds__(128) adaptive_tile_kernel(const FxTileParams p) {
    __shared__ uint32_t ambQ[4][kAdQ];
    __shared__ int ambN[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) ambN[warp] = 0;
    __syncwarp();
    const int x0 = (blockIdx.x * 128 + threadIdx.x) * 4;
    const int yb = blockIdx.y * kFxRows, img = blockIdx.z;
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;
    uint8_t *d = p.dst + (long long)img * p.dstImgStride;
    const bool active = x0 < p.w;                     // inactive lanes still take part in the queue drains
    const int xa = active ? x0 : 0;
    const bool full = p.vecOK && xa + 4 <= p.w;
    const int xl = max(xa - 1, 0), xr = min(xa + 4, p.w - 1);
    const bool dvec = full && ((((uintptr_t)p.dst | (uintptr_t)p.dstImgStride | (uintptr_t)p.dstRowStride) & 15) == 0);

    auto load_row = [&](int y, uint32_t (&px)[6]) {
        const uint8_t *row = s + (long long)min(max(y, 0), p.h - 1) * p.srcRowStride;
        if (full) {
            uint4 q = *reinterpret_cast<const uint4 *>(row + (long long)xa * 4);
            px[1] = q.x; px[2] = q.y; px[3] = q.z; px[4] = q.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) px[1 + i] = ld_nc_u32(row + (long long)min(xa + i, p.w - 1) * 4);
        }
        px[0] = ld_nc_u32(row + (long long)xl * 4);
        px[5] = ld_nc_u32(row + (long long)xr * 4);
    };
    auto hsum = [&](const uint32_t (&px)[6], uint32_t (&hrb)[4], uint32_t (&hga)[4]) {
        uint32_t rb[6], ga[6];
#pragma unroll
        for (int i = 0; i < 6; i++) { rb[i] = px[i] & 0x00FF00FFu; ga[i] = (px[i] >> 8) & 0x00FF00FFu; }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            hrb[i] = rb[i] + 2 * rb[i + 1] + rb[i + 2];
            hga[i] = ga[i] + 2 * ga[i + 1] + ga[i + 2];
        }
    };
    // integer lumas x1000 and their horizontal 1-2-1 / difference combinations for the Sobel sums
    auto lumas = [&](const uint32_t (&px)[6], int (&L)[6]) {
#pragma unroll
        for (int i = 0; i < 6; i++) L[i] = (int)__dp2a_lo(299u | (587u << 16), px[i], __dp2a_hi(114u, px[i], 0u));
    };
    const float amountF = (float)p.amount;
    const float kMagic = 12582912.0f;

    // warp-uniform: every tile of the warp strictly inside the image, rows 16-byte aligned
    if (p.fastTiles && __all_sync(0xffffffffu, active && dvec && xa >= 4 && xa + 5 <= p.w && yb >= 1 && yb + kFxRows + 1 <= p.h)) {
        const uint8_t *row = s + (long long)(yb - 1) * p.srcRowStride + (long long)xa * 4;
        uint8_t *drow = d + (long long)yb * p.dstRowStride + (long long)xa * 4;
#pragma unroll
        for (int r = 2; r < kFxRows + 2; r++) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + (long long)r * p.srcRowStride));
        // one row: 1-2-1 sums of the colour lanes, integer lumas x1000 of the six pixels and their 1-2-1 sums, raw pixels
        auto load_row_fast = [&](const uint8_t *r, uint32_t (&hrb)[4], uint32_t (&hga)[4], int (&L)[6], int (&hl)[4], uint32_t (&raw)[4]) {
            const uint4 q = ld_nc_u128(r);
            const uint32_t px[6] = {ld_nc_u32(r - 4), q.x, q.y, q.z, q.w, ld_nc_u32(r + 16)};
            uint32_t rb[6], ga[6];
#pragma unroll
            for (int i = 0; i < 6; i++) {
                rb[i] = px[i] & 0x00FF00FFu; ga[i] = __byte_perm(px[i], 0u, 0x4341u);
                L[i] = (int)__dp2a_lo(299u | (587u << 16), px[i], __dp2a_hi(114u, px[i], 0u));
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                hrb[i] = rb[i] + 2 * rb[i + 1] + rb[i + 2];
                hga[i] = ga[i] + 2 * ga[i + 1] + ga[i + 2];
                hl[i] = L[i] + 2 * L[i + 1] + L[i + 2];
                raw[i] = px[i + 1];
            }
        };
        uint32_t hP_rb[4], hP_ga[4], hC_rb[4], hC_ga[4], hN_rb[4], hN_ga[4], rawP[4], rawC[4], rawN[4];
        int lP[6], lC[6], lN[6], hlP[4], hlC[4], hlN[4];
        load_row_fast(row, hP_rb, hP_ga, lP, hlP, rawP);
        load_row_fast(row + p.srcRowStride, hC_rb, hC_ga, lC, hlC, rawC);
        row += 2 * (long long)p.srcRowStride;
#pragma unroll (kAdUnroll)
        for (int r = 0; r < kFxRows; r++, row += p.srcRowStride, drow += p.dstRowStride) {
            load_row_fast(row, hN_rb, hN_ga, lN, hlN, rawN);
            int colsum[6];
#pragma unroll
            for (int i = 0; i < 6; i++) colsum[i] = lP[i] + 2 * lC[i] + lN[i];
            uint32_t out[4];
            const uint32_t ambMask = adaptive_row_fast(hP_rb, hP_ga, hC_rb, hC_ga, hN_rb, hN_ga, hlP, hlN, colsum, rawC, amountF, out);
            *reinterpret_cast<uint4 *>(drow) = make_uint4(out[0], out[1], out[2], out[3]);
            amb_push(ambMask, xa, yb + r, 1, 0, ambQ[warp], &ambN[warp]);
            __syncwarp();
            {   // drain 32 at a time (warp-uniform decision on lane 0's view, see amb_drain)
                int n = __shfl_sync(0xffffffffu, ambN[warp], 0);
                if (n >= 32) {
                    while (n >= 32) {
                        n -= 32;
                        const uint32_t code = ambQ[warp][n + lane];
                        const int ex = (int)(code & 0xFFFFu), ey = (int)(code >> 16);
                        *reinterpret_cast<uint32_t *>(d + (long long)ey * p.dstRowStride + (long long)ex * 4) = adaptive_exact_at(s, p.srcRowStride, ex, ey, p.amount);
                    }
                    __syncwarp();
                    if (lane == 0) ambN[warp] = n;
                    __syncwarp();
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                hP_rb[i] = hC_rb[i]; hP_ga[i] = hC_ga[i]; hC_rb[i] = hN_rb[i]; hC_ga[i] = hN_ga[i];
                hlP[i] = hlC[i]; hlC[i] = hlN[i]; rawC[i] = rawN[i];
            }
#pragma unroll
            for (int i = 0; i < 6; i++) { lP[i] = lC[i]; lC[i] = lN[i]; }
        }
        __syncwarp();
        {   // leftovers
            const int n = __shfl_sync(0xffffffffu, ambN[warp], 0);
            if (lane < n) {
                const uint32_t code = ambQ[warp][lane];
                const int ex = (int)(code & 0xFFFFu), ey = (int)(code >> 16);
                *reinterpret_cast<uint32_t *>(d + (long long)ey * p.dstRowStride + (long long)ex * 4) = adaptive_exact_at(s, p.srcRowStride, ex, ey, p.amount);
            }
        }
        return;
    }

    uint32_t pPrev[6], pCur[6], pNext[6];
    uint32_t hPrevRB[4], hPrevGA[4], hCurRB[4], hCurGA[4], hNextRB[4], hNextGA[4];
    int lPrev[6], lCur[6], lNext[6];
    load_row(yb - 1, pPrev);
    load_row(yb, pCur);
    hsum(pPrev, hPrevRB, hPrevGA);
    hsum(pCur, hCurRB, hCurGA);
    lumas(pPrev, lPrev);
    lumas(pCur, lCur);
#pragma unroll 1
    for (int r = 0; r < kFxRows; r++) {
        const int y = yb + r;
        if (y >= p.h) break;   // warp-uniform
        load_row(y + 1, pNext);
        hsum(pNext, hNextRB, hNextGA);
        lumas(pNext, lNext);
        int colsum[6];
#pragma unroll
        for (int i = 0; i < 6; i++) colsum[i] = lPrev[i] + 2 * lCur[i] + lNext[i];
        uint32_t out[4];
        uint32_t ambMask = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int x = xa + i;
            const uint32_t c = pCur[1 + i];
            uint32_t res = c;
            const bool interior = x >= 1 && x < p.w - 1 && y >= 1 && y < p.h - 1;
            if (interior) {
                const uint32_t brb = ((hPrevRB[i] + 2 * hCurRB[i] + hNextRB[i] + 0x00080008u) >> 4) & 0x00FF00FFu;
                const uint32_t bga = ((hPrevGA[i] + 2 * hCurGA[i] + hNextGA[i] + 0x00080008u) >> 4) & 0x00FF00FFu;
                const int GX = colsum[i + 2] - colsum[i];
                const int GY = (lNext[i] + 2 * lNext[i + 1] + lNext[i + 2]) - (lPrev[i] + 2 * lPrev[i + 1] + lPrev[i + 2]);
                const float gxf = __int_as_float(GX + 0x4B400000) - kMagic;
                const float gyf = __int_as_float(GY + 0x4B400000) - kMagic;
                const float g2 = fmaxf(fmaf(gxf, gxf, gyf * gyf), 1e-30f);
                const float la = amountF * fminf(g2 * rsqrtf(g2) * 2.5e-6f, 1.0f);
                const float lim = 0.5f - fmaf(la, 255.0f * 7.5e-7f, 8e-5f);
                // channels as floats straight from the bytes: [byte, 0, 0x40, 0x4B] = bits of 1.5 * 2^23 + byte
                const float oR = __uint_as_float(__byte_perm(c, 0x4B400000u, 0x7650u)) - kMagic;
                const float oG = __uint_as_float(__byte_perm(c, 0x4B400000u, 0x7651u)) - kMagic;
                const float oB = __uint_as_float(__byte_perm(c, 0x4B400000u, 0x7652u)) - kMagic;
                const float bR = __uint_as_float(__byte_perm(brb, 0x4B400000u, 0x7650u)) - kMagic;
                const float bG = __uint_as_float(__byte_perm(bga, 0x4B400000u, 0x7650u)) - kMagic;
                const float bB = __uint_as_float(__byte_perm(brb, 0x4B400000u, 0x7652u)) - kMagic;
                const float o3[3] = {oR, oG, oB}, b3[3] = {bR, bG, bB};
                float t3[3];
                float worst = 0.f;
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    float v = fmaf(la, o3[ch] - b3[ch], o3[ch]);   // orig - blur is an exact small integer
                    v = fminf(fmaxf(v, 0.0f), 255.0f);             // ties at -0.5 / 255.5 cannot change the clamped result
                    t3[ch] = v + kMagic;
                    worst = fmaxf(worst, fabsf(v - (t3[ch] - kMagic)));
                }
                if (worst >= lim) ambMask |= 1u << i;
                res = pack_rgba_low_bytes(t3[0], t3[1], t3[2], c);
            }
            out[i] = res;
        }
        if (active) {
            uint8_t *drow = d + (long long)y * p.dstRowStride + (long long)xa * 4;
            if (dvec) {
                *reinterpret_cast<uint4 *>(drow) = make_uint4(out[0], out[1], out[2], out[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (xa + i < p.w) *reinterpret_cast<uint32_t *>(drow + i * 4) = out[i];
            }
            if (xa + 4 > p.w) ambMask &= (1u << (p.w - xa)) - 1u;
            amb_push(ambMask, xa, y, 1, 0, ambQ[warp], &ambN[warp]);
        }
        __syncwarp();
        {   // drain 32 at a time (warp-uniform decision on lane 0's view, see amb_drain)
            int n = __shfl_sync(0xffffffffu, ambN[warp], 0);
            if (n >= 32) {
                while (n >= 32) {
                    n -= 32;
                    const uint32_t code = ambQ[warp][n + lane];
                    const int ex = (int)(code & 0xFFFFu), ey = (int)(code >> 16);
                    *reinterpret_cast<uint32_t *>(d + (long long)ey * p.dstRowStride + (long long)ex * 4) = adaptive_exact_at(s, p.srcRowStride, ex, ey, p.amount);
                }
                __syncwarp();
                if (lane == 0) ambN[warp] = n;
                __syncwarp();
            }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) { pPrev[i] = pCur[i]; pCur[i] = pNext[i]; lPrev[i] = lCur[i]; lCur[i] = lNext[i]; }
#pragma unroll
        for (int i = 0; i < 4; i++) { hPrevRB[i] = hCurRB[i]; hPrevGA[i] = hCurGA[i]; hCurRB[i] = hNextRB[i]; hCurGA[i] = hNextGA[i]; }
    }
    __syncwarp();
    {   // leftovers
        const int n = __shfl_sync(0xffffffffu, ambN[warp], 0);
        if (lane < n) {
            const uint32_t code = ambQ[warp][lane];
            const int ex = (int)(code & 0xFFFFu), ey = (int)(code >> 16);
            *reinterpret_cast<uint32_t *>(d + (long long)ey * p.dstRowStride + (long long)ex * 4) = adaptive_exact_at(s, p.srcRowStride, ex, ey, p.amount);
        }
    }
}

}  // namespace

int launch_gaussian_blur(cudaStream_t s, const uint8_t *src, uint8_t *dst, long long imgStride,
                         int rowStride, int w, int h, int n, const double *kernel_dev,
                         const float *kernel32_dev, const float *kernel32_host, int radius, double wabs, uint8_t *tmp,
                         long long tmpImgStride, int tmpRowStride) {
    if (n <= 0 || w <= 0 || h <= 0) return FB_OK;
    BlurParams p;
    p.w = w; p.h = h; p.radius = radius;
    p.kernel = kernel_dev; p.kernel32 = kernel32_dev;
    for (int k = 0; k < 17; k++) p.w32[k] = (k <= 2 * radius && radius <= 8) ? kernel32_host[k] : 0.f;
    // FP32 error bound of the tap sum: FMA k rounds a partial sum <= 255 * P_k, P_k = sum_{s<=k} |w_s| in tap order (the
    // order both fast kernels accumulate in), and the FP32 weights differ from the FP64 ones by <= 2^-24 relative
    // (255 * 2^-24 * sum|w| in total): eps = 255 * 2^-24 * (sum|w| + sum_k P_k), 10 % margin.  [Round 1 used
    // (taps + 1) * 255 * 2^-24 * 1.25, twice as wide for a Gaussian: twice as many outputs on the exact path.]
    // A caller-supplied kernel that is not a convex combination arrives with wabs = 1e9 (api.cu): exact path only.
    double sabs = 0.0, psum = 0.0;
    for (int k = 0; k <= 2 * radius; k++) { sabs += fabs((double)kernel32_host[k]); psum += sabs; }
    const bool sane = (wabs == wabs && wabs < 1e6 && sabs == sabs);
    p.eps = sane ? (float)(255.0 * 5.9604644775390625e-08 * (sabs + psum) * 1.10) : 1e6f;
    p.exactOnly = (p.eps >= 0.25f) ? 1 : 0;  // absurdly long (or huge-gain) kernels: no useful fast path
    dim3 grid((w + 255) / 256, h, n);
    // horizontal: src → tmp
    p.src = src; p.srcImgStride = imgStride; p.srcRowStride = rowStride;
    p.alpha = src; p.alphaImgStride = imgStride; p.alphaRowStride = rowStride;
    p.dst = tmp; p.dstImgStride = tmpImgStride; p.dstRowStride = tmpRowStride;
    const bool fastOK = !p.exactOnly && getenv("FB_BLUR_GENERIC") == nullptr && w <= 65535 && h <= 65535;  // the exact queue packs (x, y) in 16+16 bits
    if (!(fastOK && launch_blur_fast_any(s, p, n, false))) blur_pass_kernel<false><<<grid, 256, 0, s>>>(p);
    // vertical: tmp → dst, alpha from the original
    p.src = tmp; p.srcImgStride = tmpImgStride; p.srcRowStride = tmpRowStride;
    p.dst = dst; p.dstImgStride = imgStride; p.dstRowStride = rowStride;
    if (!(fastOK && launch_blur_fast_any(s, p, n, true))) blur_pass_kernel<true><<<grid, 256, 0, s>>>(p);
    FB_LAUNCHED(2);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

static int launch_fx(cudaStream_t s, const uint8_t *src, uint8_t *dst, long long imgStride, int rowStride,
                     int w, int h, int n, long long dstImgStride, int dstRowStride, double amount, int mode) {
    if (n <= 0 || w <= 0 || h <= 0) return FB_OK;
    FxParams p;
    p.src = src; p.dst = dst;
    p.srcImgStride = imgStride; p.dstImgStride = dstImgStride;
    p.srcRowStride = rowStride; p.dstRowStride = dstRowStride;
    p.w = w; p.h = h; p.amount = amount; p.mode = mode;
    if (getenv("FB_FX_GENERIC") == nullptr) {
        FxTileParams t;
        t.src = src; t.dst = dst;
        t.srcImgStride = imgStride; t.dstImgStride = dstImgStride;
        t.srcRowStride = rowStride; t.dstRowStride = dstRowStride;
        t.w = w; t.h = h; t.amount = amount; t.A = 0; t.half = 0;
        t.vecOK = (((uintptr_t)src | (uintptr_t)imgStride | (uintptr_t)rowStride) & 15) == 0;
        t.fastTiles = getenv("FB_FX_NOFAST") == nullptr ? 1 : 0;
        dim3 tgrid((w + 511) / 512, (h + kFxRows - 1) / kFxRows, n);
        int k = -1;
        if (mode == 1) {  // Sharpen: is amount * 2^k an integer for a small k?  (orig<<k) + A*diff must fit in int32
            for (int kk = 0; kk <= 12 && k < 0; kk++) {
                double scaled = amount * (double)(1 << kk);
                if (scaled == (double)(long long)scaled && scaled < 32768.0) { k = kk; t.A = (int)scaled; t.half = kk ? 1 << (kk - 1) : 0; }
            }
        }
        if (mode == 0) fx_tile_kernel<0, -1><<<tgrid, 128, 0, s>>>(t);
        else if (mode == 2 && w <= 65535 && h <= 65535 && getenv("FB_ADAPTIVE_OLD") == nullptr) adaptive_tile_kernel<<<tgrid, 128, 0, s>>>(t);
        else if (mode == 2) fx_tile_kernel<2, -1><<<tgrid, 128, 0, s>>>(t);
        else if (k == 0) fx_tile_kernel<1, 0><<<tgrid, 128, 0, s>>>(t);
        else if (k == 1) fx_tile_kernel<1, 1><<<tgrid, 128, 0, s>>>(t);
        else if (k == 2) fx_tile_kernel<1, 2><<<tgrid, 128, 0, s>>>(t);
        else if (k == 3) fx_tile_kernel<1, 3><<<tgrid, 128, 0, s>>>(t);
        else if (k == 4) fx_tile_kernel<1, 4><<<tgrid, 128, 0, s>>>(t);
        else fx_tile_kernel<1, -1><<<tgrid, 128, 0, s>>>(t);
    } else {
        dim3 grid((w + 255) / 256, h, n);
        sharpen_kernel<<<grid, 256, 0, s>>>(p);
    }
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

int launch_blur3x3(cudaStream_t s, const uint8_t *src, uint8_t *dst, long long imgStride, int rowStride,
                   int w, int h, int n, long long dstImgStride, int dstRowStride) {
    return launch_fx(s, src, dst, imgStride, rowStride, w, h, n, dstImgStride, dstRowStride, 0.0, 0);
}

int launch_sharpen(cudaStream_t s, const uint8_t *src, uint8_t *dst, long long imgStride, int rowStride,
                   int w, int h, int n, long long dstImgStride, int dstRowStride, double amount, int adaptive) {
    return launch_fx(s, src, dst, imgStride, rowStride, w, h, n, dstImgStride, dstRowStride, amount,
                     adaptive ? 2 : 1);
}

}  // namespace fb
