// resize.cu — K7: two-pass Lanczos-3 resampler with premultiplied alpha, bit-exact with
// resizeH / resizeV (resize.go:77-161).
//
// Per destination pixel the reference accumulates, over the taps of that destination index in
// ascending source order:  aw = alpha*w;  r += R*aw;  g += G*aw;  b += B*aw;  a += aw  — all binary64,
// unfused — then, if a > 0.5, writes clampF(r * (1/a)) ... clampF(a), else leaves the pixel zero.
// The horizontal pass result is rounded to uint8 before the vertical pass (resize.go:51-52).
// The weight tables (CSR: start/index/weight) come from the host side of the ABI.
//
// Mapping: one thread per destination pixel; a warp covers 32 adjacent destination columns, so in
// the vertical pass every tap is one coalesced 128-byte row segment, and in the horizontal pass the
// warp's taps cover one contiguous span of the source row that stays in L1.
#include "common.cuh"

#include <stdlib.h>

namespace fb {

namespace {

struct ResizeParams {
    const uint8_t *src;
    uint8_t *dst;
    long long srcImgStride, dstImgStride;
    int srcRowStride, dstRowStride;
    int outW, outH;          // dims of the destination of THIS pass
    const int *start;        // CSR over the resampled axis
    const int *index;
    const double *weight;
    const float *weight32;   // the same weights rounded to FP32 (fast path)
    float Er, Ea;            // FP32 error bounds of the colour sums and of the alpha sum (see launch_pass)
    // horizontal fast path: taps of destination d are the contiguous pixels first[d].. ; they are read as
    // 16-byte groups starting at first[d] & ~3, with FP32 weights zero-padded to whole groups and stored
    // transposed, wpadT[(q * n + d) * 4 + j], so that a warp's weight loads are contiguous.
    const int *first;
    const float *wpadT;
    int groups;              // max groups per destination
    int srcW;
    int vecOK;
};

__device__ __forceinline__ void finish_px(double r, double g, double b, double a, uint8_t *d) {
    uint32_t out = 0u;  // resize.go:107-113 — pixels with a <= 0.5 stay zero (fresh image)
    if (a > 0.5) {
        double inv = __ddiv_rn(1.0, a);
        out = clampf_dev(__dmul_rn(r, inv)) | (clampf_dev(__dmul_rn(g, inv)) << 8) |
              (clampf_dev(__dmul_rn(b, inv)) << 16) | (clampf_dev(a) << 24);
    }
    *reinterpret_cast<uint32_t *>(d) = out;
}

template <bool VERTICAL>
__global__ void __launch_bounds__(256) resize_pass_kernel(const ResizeParams p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, img = blockIdx.z;
    if (x >= p.outW) return;
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;
    const int d = VERTICAL ? y : x;
    const int t0 = __ldg(p.start + d), t1 = __ldg(p.start + d + 1);
    double r = 0.0, g = 0.0, b = 0.0, a = 0.0;
    for (int t = t0; t < t1; t++) {
        const int si = __ldg(p.index + t);
        const double w = __ldg(p.weight + t);
        const uint32_t v = VERTICAL ? ld_nc_u32(s + (long long)si * p.srcRowStride + (long long)x * 4)
                                    : ld_nc_u32(s + (long long)y * p.srcRowStride + (long long)si * 4);
        const double aw = __dmul_rn((double)(v >> 24), w);          // resize.go:99
        r = __dadd_rn(r, __dmul_rn((double)(v & 0xFF), aw));         // :100
        g = __dadd_rn(g, __dmul_rn((double)((v >> 8) & 0xFF), aw));  // :101
        b = __dadd_rn(b, __dmul_rn((double)((v >> 16) & 0xFF), aw)); // :102
        a = __dadd_rn(a, aw);                                        // :103
    }
    finish_px(r, g, b, a, p.dst + (long long)img * p.dstImgStride + (long long)y * p.dstRowStride + (long long)x * 4);
}

__device__ __forceinline__ float byte_f(uint32_t px, int k) {  // byte k of px as float, without I2F
    return __uint_as_float(__byte_perm(px, 0x4B000000u, 0x7540u | (uint32_t)k)) - 8388608.0f;
}

// Round v in [0,255] to nearest; `amb` when within eps of a tie (the exact path then decides).
__device__ __forceinline__ uint32_t round_flag255(float v, float eps, bool &amb) {
    v = fminf(fmaxf(v, 0.f), 255.f);  // ties at -0.5 and 255.5 cannot change the clamped result
    float t = v + 12582912.0f;
    float rounded = t - 12582912.0f;
    amb = amb || (fabsf(v - rounded) >= 0.5f - eps);
    return __float_as_uint(t) & 0x1FFu;
}

// Fast pass: the premultiplied sums are accumulated with FP32 FMAs.  |r32 - r64| <= Er and
// |a32 - a64| <= Ea (bounds from the tap count and max sum|w| of the table), so the quotient is within
// eps = (Er + |v|*Ea)/a + |v|*2^-21 of the reference value; a pixel whose channel (or alpha, or the
// a > 0.5 gate) lies within that distance of a decision boundary is recomputed in the exact FP64 order.
// Ambiguous pixels are compacted per block through shared memory so the exact path runs with full
// warps instead of one divergent lane per warp.
__device__ __forceinline__ void tap_fp32(uint32_t v, float w, float &r, float &g, float &b, float &a) {
    const float aw = byte_f(v, 3) * w;
    r = fmaf(byte_f(v, 0), aw, r);
    g = fmaf(byte_f(v, 1), aw, g);
    b = fmaf(byte_f(v, 2), aw, b);
    a += aw;
}

// Returns the packed fast result, or sets amb.
__device__ __forceinline__ uint32_t finish_fp32(float r, float g, float b, float a, float Er, float Ea, bool &amb) {
    amb = fabsf(a - 0.5f) <= Ea;  // the a > 0.5 gate itself (resize.go:107)
    uint32_t out = 0u;
    if (!amb && a > 0.5f) {
        const float inv = __frcp_rn(a);
        const float vr = r * inv, vg = g * inv, vb = b * inv;
        const float vmax = fmaxf(fmaxf(fabsf(vr), fabsf(vg)), fabsf(vb));
        const float eps = (Er + vmax * Ea) * inv + vmax * 4.76837158203125e-07f;
        amb = !(eps < 0.25f);
        const uint32_t cr = round_flag255(vr, eps, amb), cg = round_flag255(vg, eps, amb), cb = round_flag255(vb, eps, amb);
        const uint32_t ca = round_flag255(a, Ea, amb);
        out = cr | (cg << 8) | (cb << 16) | (ca << 24);
    }
    return out;
}

// Exact FP64 sequence of the reference for one destination pixel (resize.go:93-113).  Taps are fetched eight at
// a time (independent loads in flight) and bytes become doubles through the 2^52 mantissa trick, so the rare
// exact path is neither a chain of dependent global loads nor a queue on the conversion unit.
__device__ __forceinline__ double byte_to_double(uint32_t b) {
    return __hiloint2double(0x43300000, (int)b) - 4503599627370496.0;  // (2^52 + b) - 2^52, exact
}

template <bool VERTICAL>
__device__ __forceinline__ void exact_px(const ResizeParams &p, const uint8_t *s, int x, int y, uint8_t *dpx) {
    const int d = VERTICAL ? y : x;
    const int t0 = __ldg(p.start + d), t1 = __ldg(p.start + d + 1);
    double r2 = 0.0, g2 = 0.0, b2 = 0.0, a2 = 0.0;
    for (int tb = t0; tb < t1; tb += 8) {
        uint32_t v[8];
        double w[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int t = min(tb + k, t1 - 1);
            const int si = __ldg(p.index + t);
            w[k] = __ldg(p.weight + t);
            v[k] = VERTICAL ? __ldg(reinterpret_cast<const uint32_t *>(s + (long long)si * p.srcRowStride + (long long)x * 4))
                            : __ldg(reinterpret_cast<const uint32_t *>(s + (long long)y * p.srcRowStride + (long long)si * 4));
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (tb + k < t1) {
                const double aw = __dmul_rn(byte_to_double(v[k] >> 24), w[k]);
                r2 = __dadd_rn(r2, __dmul_rn(byte_to_double(v[k] & 0xFF), aw));
                g2 = __dadd_rn(g2, __dmul_rn(byte_to_double((v[k] >> 8) & 0xFF), aw));
                b2 = __dadd_rn(b2, __dmul_rn(byte_to_double((v[k] >> 16) & 0xFF), aw));
                a2 = __dadd_rn(a2, aw);
            }
        }
    }
    finish_px(r2, g2, b2, a2, dpx);
}

// FP32 sums of one destination pixel through the general (CSR / grouped) tables.
template <bool VERTICAL, bool GROUPED>
__device__ __forceinline__ void general_sums(const ResizeParams &p, const uint8_t *s, int x, int y,
                                             float &r, float &g, float &b, float &a) {
    r = g = b = a = 0.f;
    if (GROUPED) {  // horizontal: 128-bit pixel groups + transposed padded weights
        const int first = __ldg(p.first + x);
        const int g0 = first & ~3;
        const int ng = ((first + (__ldg(p.start + x + 1) - __ldg(p.start + x)) - 1) >> 2) - (g0 >> 2) + 1;
        const uint8_t *row = s + (long long)y * p.srcRowStride;
        for (int q = 0; q < ng; q++) {
            const int px0 = g0 + 4 * q;
            const float4 w4 = __ldg(reinterpret_cast<const float4 *>(p.wpadT) + (size_t)q * p.outW + x);
            uint32_t v[4];
            if (p.vecOK && px0 + 4 <= p.srcW) {
                uint4 t = __ldg(reinterpret_cast<const uint4 *>(row + (long long)px0 * 4));
                v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++)  // pixels past the row end carry zero weights
                    v[j] = (px0 + j < p.srcW) ? __ldg(reinterpret_cast<const uint32_t *>(row + (long long)(px0 + j) * 4)) : 0u;
            }
            tap_fp32(v[0], w4.x, r, g, b, a);
            tap_fp32(v[1], w4.y, r, g, b, a);
            tap_fp32(v[2], w4.z, r, g, b, a);
            tap_fp32(v[3], w4.w, r, g, b, a);
        }
    } else {
        const int d = VERTICAL ? y : x;
        const int t0 = __ldg(p.start + d), t1 = __ldg(p.start + d + 1);
        for (int t = t0; t < t1; t++) {
            const int si = __ldg(p.index + t);
            const float w = __ldg(p.weight32 + t);
            const uint32_t v = VERTICAL ? __ldg(reinterpret_cast<const uint32_t *>(s + (long long)si * p.srcRowStride + (long long)x * 4))
                                        : __ldg(reinterpret_cast<const uint32_t *>(s + (long long)y * p.srcRowStride + (long long)si * 4));
            tap_fp32(v, w, r, g, b, a);
        }
    }
}

template <bool VERTICAL, bool GROUPED>
__global__ void __launch_bounds__(256) resize_pass_fast_kernel(const ResizeParams p) {
    __shared__ int nAmb;
    __shared__ unsigned short ambList[256];
    if (threadIdx.x == 0) nAmb = 0;
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, img = blockIdx.z;
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;
    uint8_t *drow = p.dst + (long long)img * p.dstImgStride + (long long)y * p.dstRowStride;
    if (x < p.outW) {
        float r, g, b, a;
        general_sums<VERTICAL, GROUPED>(p, s, x, y, r, g, b, a);
        bool amb;
        const uint32_t out = finish_fp32(r, g, b, a, p.Er, p.Ea, amb);
        if (amb) ambList[atomicAdd(&nAmb, 1)] = (unsigned short)threadIdx.x;
        else *reinterpret_cast<uint32_t *>(drow + (long long)x * 4) = out;
    }
    __syncthreads();
    const int n = nAmb;
    if ((int)threadIdx.x < n) {
        const int xa = blockIdx.x * blockDim.x + ambList[threadIdx.x];
        exact_px<VERTICAL>(p, s, xa, y, drow + (long long)xa * 4);
    }
}

// ------------------------------------------------------------------------------------------------
// Integer-ratio kernel (srcSize == R * dstSize; config 4 is R = 4).  Every interior destination index d
// then has the same T taps at source R*d + off with the SAME weights (verified on the host), so a thread
// can produce kOut adjacent outputs from one window of T + (kOut-1)*R pixels with compile-time tap
// indices: every byte is converted once and feeds up to kOut outputs (FFMA2 on the (R,G) and (B,A) channel
// pairs with the tap weight as a broadcast scalar).  Windows that are fully opaque (alpha == 255 everywhere) skip the
// premultiplication: v = sum(R*w) / sum(w).  Edge outputs and windows with any translucent pixel take the
// general path per output.  Ambiguous results go to the block-compacted exact FP64 path as above.
// ------------------------------------------------------------------------------------------------
constexpr int kOut = 4;

struct IntRatioParams {
    ResizeParams base;
    int off;        // first tap of destination d is R*d + off
    int dLo, dHi;   // destinations in [dLo, dHi) are interior (identical weights, no clipping)
    float wsum;     // FP32 sum of the weights
    float w[28];    // the T shared weights (T <= 24 used)
    float wn[28];   // w / sum(w): a fully opaque window is v = sum(R * wn), no premultiply and no division
    float Eo;       // |FP32 value - reference value| of that sum (detect_int_ratio, api.cu)
    uint32_t opaqueA;  // alpha byte << 24 such a window produces; 0: shortcut disabled
    int wdExact;       // every interior destination has bit-identical binary64 weights: the exact path can use wd[] below
    double wd[24];     // ... instead of the CSR index and weight rows (no dependent index -> pixel loads)
};

#ifndef FB_LZ_ROWS
#define FB_LZ_ROWS 8
#endif
constexpr int kRowsPerBlock = FB_LZ_ROWS;  // horizontal pass: rows one block walks (double-buffered staging); <= 8 (queue row code is 3 bits)

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// max(min(round(v), 255), 0) from t = v + 1.5 * 2^23 (|v| < 2^22): the low 23 bits of t are 2^22 + round(v), so one
// VIADDMNMX.RELU on the bit pattern subtracts the magic, clamps above and below.
__device__ __forceinline__ uint32_t clamp255_from_magic(float t) {
    return (uint32_t)__viaddmin_s32_relu(__float_as_int(t), -0x4B400000, 255);
}

// finish_fp32 for the integer-ratio window kernels, same decisions with ~30 instead of 50 instructions: the reciprocal is the
// MUFU approximation (1 ulp, covered by the 6.0e-7 = 2^-21 + 2^-23 term of the bound), the magic-number rounding runs as
// packed FADD2 on the (R,G) and (B,A) pairs, the clamp is one VIADDMNMX.RELU per channel on the bit pattern and the pack
// three PRMTs.  Values outside [0, 255] are flagged at the same rate as inside (harmless: the exact path clamps too).
__device__ __forceinline__ uint32_t finish_fp32_lean(float2 rg, float2 ba, float Er, float Ea, bool &amb) {
    const float a = ba.y;
    amb = fabsf(a - 0.5f) <= Ea;  // the a > 0.5 gate itself (resize.go:107)
    uint32_t out = 0u;
    if (!amb && a > 0.5f) {
        float inv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(a));
        const float2 vrg = __fmul2_rn(rg, make_float2(inv, inv));
        const float2 vba = make_float2(ba.x * inv, a);
        const float vmax = fmaxf(fmaxf(fabsf(vrg.x), fabsf(vrg.y)), fabsf(vba.x));
        const float eps = fmaf(fmaf(vmax, Ea, Er), inv, vmax * 6.0e-7f);
        const float2 magic = make_float2(12582912.0f, 12582912.0f), nmagic = make_float2(-12582912.0f, -12582912.0f);
        const float2 neg1 = make_float2(-1.0f, -1.0f);
        const float2 t1 = __fadd2_rn(vrg, magic), t2 = __fadd2_rn(vba, magic);
        const float2 d1 = __ffma2_rn(__fadd2_rn(t1, nmagic), neg1, vrg), d2 = __ffma2_rn(__fadd2_rn(t2, nmagic), neg1, vba);
        amb = !(eps < 0.25f) || fmaxf(fmaxf(fabsf(d1.x), fabsf(d1.y)), fabsf(d2.x)) >= 0.5f - eps || fabsf(d2.y) >= 0.5f - Ea;
        const uint32_t x = __byte_perm(clamp255_from_magic(t1.x), clamp255_from_magic(t1.y), 0x0040);
        const uint32_t z = __byte_perm(clamp255_from_magic(t2.x), clamp255_from_magic(t2.y), 0x0040);
        out = __byte_perm(x, z, 0x5410);
    }
    return out;
}

#ifndef FB_LZ_I2F
#define FB_LZ_I2F 1   // 1: the B (and alpha) bytes become floats through I2F.U8 on the otherwise idle XU pipe (16 lanes/clk/SM) instead of
                     // PRMT + FADD: 0.550 against 0.568 ms per 8 opaque 8K images; 2: R and G too — the XU pipe saturates, 0.651 ms
#endif
// KO interior outputs from one window of packed pixels (compile-time tap indices); the window starts at raw[LEAD].
template <int R, int T, int LEAD, int NRAW, int KO = kOut>
__device__ __forceinline__ void int_ratio_window(const uint32_t (&raw)[NRAW], const IntRatioParams &q,
                                                 uint32_t (&outv)[KO], bool (&ambv)[KO]) {
    constexpr int NIN = T + (KO - 1) * R;
    static_assert(NRAW >= NIN + LEAD && KO % 2 == 0, "window does not fit");
    const ResizeParams &p = q.base;
    uint32_t andA = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < NIN; i++) andA &= raw[LEAD + i];
    // Channel-paired accumulators: (R, G) of one output share an FFMA2 whose weight is a broadcast scalar
    // (w[t] straight from the parameter bank) and B (and alpha) ride in a second accumulator.  [An earlier
    // version paired two OUTPUTS per FFMA2, which needs the weight pairs (w[t], w[t-R]) in registers: ptxas
    // rebuilt every pair with two MOVs per FFMA2 — 3x the instructions, profiles/r1b_ncu_summaries_all_kernels.txt.]
    const float2 kMagic2 = make_float2(-8388608.0f, -8388608.0f);
    if (q.opaqueA != 0u && (andA >> 24) == 0xFFu) {
        // Fully opaque window: the reference's r/a = sum(R*255*w) / sum(255*w) is sum(R * w/W) up to 1e-13, so the
        // normalised weights give the value directly: no premultiply, no reciprocal, a constant alpha byte and a
        // constant error bound Eo (detect_int_ratio).  Round + clamp + pack: magic-number rounding as packed FADD2
        // on the (R,G) pair and on the B values of two outputs, one VIADDMNMX.RELU per channel (Lanczos overshoots,
        // so the clamp is real), three PRMTs.  [The general finish_fp32 — reciprocal, per-pixel bound, float clamps —
        // was 55 of the 168 instructions per output, profiles/r2s2_lanczos.]
        float2 accRG[KO], accB[KO / 2];
#pragma unroll
        for (int j = 0; j < KO; j++) accRG[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < KO / 2; m++) accB[m] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NIN; i++) {
            const uint32_t px = raw[LEAD + i];
#if FB_LZ_I2F >= 2
            const float2 rg = make_float2((float)(px & 0xFFu), (float)((px >> 8) & 0xFFu));
#else
            const float2 rg = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7540u)),
                                                     __uint_as_float(__byte_perm(px, 0x4B000000u, 0x7541u))), kMagic2);
#endif
#if FB_LZ_I2F
            const float bl = (float)((px >> 16) & 0xFFu);   // I2F.U8 on the XU pipe instead of PRMT + FADD
#else
            const float bl = byte_f(px, 2);
#endif
#pragma unroll
            for (int j = 0; j < KO; j++) {
                const int t = i - j * R;  // tap of input i for output j
                if (t >= 0 && t < T) {
                    const float w = q.wn[t >= 0 && t < T ? t : 0];
                    accRG[j] = __ffma2_rn(rg, make_float2(w, w), accRG[j]);
                    if (j & 1) accB[j / 2].y = fmaf(bl, w, accB[j / 2].y);
                    else accB[j / 2].x = fmaf(bl, w, accB[j / 2].x);
                }
            }
        }
        const float2 magic = make_float2(12582912.0f, 12582912.0f), nmagic = make_float2(-12582912.0f, -12582912.0f);
        const float2 neg1 = make_float2(-1.0f, -1.0f);
        const float lim = 0.5f - q.Eo;
#pragma unroll
        for (int m = 0; m < KO / 2; m++) {
            float2 t[3], d[3];   // (R,G) of output 2m, (R,G) of output 2m+1, B of both
            const float2 v[3] = {accRG[2 * m], accRG[2 * m + 1], accB[m]};
#pragma unroll
            for (int c = 0; c < 3; c++) {
                t[c] = __fadd2_rn(v[c], magic);
                const float2 r = __fadd2_rn(t[c], nmagic);
                d[c] = __ffma2_rn(r, neg1, v[c]);   // v - round(v), exact
            }
            ambv[2 * m] = fmaxf(fmaxf(fabsf(d[0].x), fabsf(d[0].y)), fabsf(d[2].x)) >= lim;
            ambv[2 * m + 1] = fmaxf(fmaxf(fabsf(d[1].x), fabsf(d[1].y)), fabsf(d[2].y)) >= lim;
            const uint32_t x0 = __byte_perm(clamp255_from_magic(t[0].x), clamp255_from_magic(t[0].y), 0x0040);
            const uint32_t x1 = __byte_perm(clamp255_from_magic(t[1].x), clamp255_from_magic(t[1].y), 0x0040);
            outv[2 * m] = __byte_perm(x0, __byte_perm(clamp255_from_magic(t[2].x), q.opaqueA, 0x7000), 0x7610);
            outv[2 * m + 1] = __byte_perm(x1, __byte_perm(clamp255_from_magic(t[2].y), q.opaqueA, 0x7000), 0x7610);
        }
    } else {  // translucent window (or shortcut disabled): premultiply once per pixel (R*alpha is an exact integer), 4 sums per output
        float2 accRG[KO], accBA[KO];
#pragma unroll
        for (int j = 0; j < KO; j++) accRG[j] = accBA[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NIN; i++) {
            const uint32_t px = raw[LEAD + i];
            const float fa = byte_f(px, 3), fb = byte_f(px, 2);   // (I2F for these two was measured slower here: 0.819 against 0.803 ms)
            const float2 rg = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7540u)),
                                                     __uint_as_float(__byte_perm(px, 0x4B000000u, 0x7541u))), kMagic2);
            const float2 prg = __fmul2_rn(rg, make_float2(fa, fa));
            const float2 pba = make_float2(fb * fa, fa);
#pragma unroll
            for (int j = 0; j < KO; j++) {
                const int t = i - j * R;
                if (t >= 0 && t < T) {
                    const float w = q.w[t >= 0 && t < T ? t : 0];
                    accRG[j] = __ffma2_rn(prg, make_float2(w, w), accRG[j]);
                    accBA[j] = __ffma2_rn(pba, make_float2(w, w), accBA[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < KO; j++) outv[j] = finish_fp32_lean(accRG[j], accBA[j], p.Er, p.Ea, ambv[j]);
    }
}

// Vertical pass: thread = column x, kOut adjacent rows; the window is NIN independent global loads (each a
// coalesced 128-byte row segment per warp).  Edge rows (clipped / renormalised taps) take the general tables.
template <int R, int T>
__global__ void __launch_bounds__(128) resize_v_int_ratio_kernel(const IntRatioParams q) {
    constexpr int NIN = T + (kOut - 1) * R;
    const ResizeParams &p = q.base;
    __shared__ int nAmb;
    __shared__ unsigned short ambList[128 * kOut];
    if (threadIdx.x == 0) nAmb = 0;
    const int img = blockIdx.z;
    const int x0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * kOut;
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;
    uint8_t *dimg = p.dst + (long long)img * p.dstImgStride;
    __syncthreads();
    if (x0 < p.outW && y0 < p.outH) {
        uint32_t outv[kOut];
        bool ambv[kOut];
        if (y0 >= q.dLo && y0 + kOut <= q.dHi) {
            uint32_t raw[NIN];
            const int s0 = R * y0 + q.off;  // first source row of the window
#pragma unroll
            for (int i = 0; i < NIN; i++)
                raw[i] = __ldg(reinterpret_cast<const uint32_t *>(s + (long long)(s0 + i) * p.srcRowStride + (long long)x0 * 4));
            int_ratio_window<R, T, 0, NIN>(raw, q, outv, ambv);
        } else {
#pragma unroll
            for (int j = 0; j < kOut; j++) {
                ambv[j] = false;
                outv[j] = 0u;
                if (y0 + j < p.outH) {
                    float r, g, b, a;
                    general_sums<true, false>(p, s, x0, y0 + j, r, g, b, a);
                    outv[j] = finish_fp32(r, g, b, a, p.Er, p.Ea, ambv[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kOut; j++) {
            if (y0 + j < p.outH) {
                if (ambv[j]) ambList[atomicAdd(&nAmb, 1)] = (unsigned short)(threadIdx.x * kOut + j);
                else *reinterpret_cast<uint32_t *>(dimg + (long long)(y0 + j) * p.dstRowStride + (long long)x0 * 4) = outv[j];
            }
        }
    }
    __syncthreads();
    const int n = nAmb;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int lt = ambList[e] / kOut, j = ambList[e] % kOut;
        const int ox = blockIdx.x * blockDim.x + lt, oy = y0 + j;
        exact_px<true>(p, s, ox, oy, dimg + (long long)oy * p.dstRowStride + (long long)ox * 4);
    }
}

// Horizontal pass: thread = kOut adjacent output columns; a block walks kRowsPerBlock rows of a 512-output column
// segment.  Row y+1 is staged into the second shared-memory buffer with cp.async while row y is evaluated (the
// first version staged one row with a load->store loop and paid the DRAM latency ~8 times per block: half of
// its stall samples, profiles/r1b_ncu_summaries_all_kernels.txt).  Ambiguous outputs AND the few edge outputs (clipped taps) are only
// queued; the queue is drained by all 128 threads after the last row (or when it could overflow), so the
// FP64 path runs with packed warps and the row loop has a single barrier.
constexpr int kAmbCap = 2048;

template <int R, int T>
__global__ void __launch_bounds__(128) resize_h_int_ratio_kernel(const IntRatioParams q) {
    constexpr int NIN = T + (kOut - 1) * R;          // window of one thread
    constexpr int SPAN = R * kOut * 128 + T - R;      // source pixels one block's row segment needs
    constexpr int CHUNKS = (SPAN + 15) / 16 + 2;
    constexpr int CHB = 80;                           // 16-px chunk + 16 B pad: conflict-free LDS.128 at 1 chunk / thread
    constexpr int STAGEB = CHUNKS * CHB;
    constexpr int NLD = (SPAN + 1 + 255) / 256;       // 64-bit staging copies per thread
    static_assert(R * kOut == 16, "staging assumes one 16-px chunk per thread");
    const ResizeParams &p = q.base;
    __shared__ int nAmb, flush[2];
    __shared__ unsigned short ambList[kAmbCap];       // (row - yFirst) * 512 + thread * kOut + j
    __shared__ __align__(16) uint8_t stage[2 * STAGEB];
    if (threadIdx.x == 0) { nAmb = 0; flush[0] = flush[1] = 0; }
    const int img = blockIdx.z;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * kOut;
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;
    uint8_t *dimg = p.dst + (long long)img * p.dstImgStride;
    const int yFirst = blockIdx.y * kRowsPerBlock;
    const int yLast = min(yFirst + kRowsPerBlock, p.outH);  // exclusive
    // Staging pixel u <-> source pixel sBase + u (sBase is even: R*kOut == 16 and even off); zero outside the row.
    const int sBase = R * (blockIdx.x * blockDim.x * kOut) + q.off;
    const bool interiorSpan = sBase >= 0 && sBase + SPAN + 2 <= p.srcW &&
                              (((uintptr_t)s + (long long)sBase * 4) & 7) == 0 && (p.srcRowStride & 7) == 0;
    auto stage_row = [&](int y, uint8_t *buf) {
        const uint8_t *row = s + (long long)y * p.srcRowStride;
        if (interiorSpan) {   // block-uniform: every copy is in range and 8-byte aligned
#pragma unroll
            for (int k = 0; k < NLD; k++) {
                const int u = threadIdx.x * 2 + k * 256;
                if (k < NLD - 1 || u < SPAN + 1) cp_async8(buf + (u >> 4) * CHB + (u & 15) * 4, row + (long long)(sBase + u) * 4);
            }
        } else {
            // edge blocks (the first and the last of a row — half of all blocks at 7680 -> 1920): pairs that lie
            // inside the row still go through cp.async; only the few pairs that straddle or leave the row are
            // loaded by hand.  [A plain load -> store loop here cost 25 % of the kernel's stall samples.]
            const bool al8 = (((uintptr_t)row + (long long)sBase * 4) & 7) == 0;
#pragma unroll
            for (int k = 0; k < NLD; k++) {
                const int u = threadIdx.x * 2 + k * 256;
                if (k < NLD - 1 || u < SPAN + 1) {
                    const int sx = sBase + u;
                    uint8_t *dstp = buf + (u >> 4) * CHB + (u & 15) * 4;
                    if (al8 && sx >= 0 && sx + 1 < p.srcW) {
                        cp_async8(dstp, row + (long long)sx * 4);
                    } else {
                        uint2 v = make_uint2(0u, 0u);
                        if (sx >= 0 && sx < p.srcW) v.x = __ldg(reinterpret_cast<const uint32_t *>(row + (long long)sx * 4));
                        if (sx + 1 >= 0 && sx + 1 < p.srcW) v.y = __ldg(reinterpret_cast<const uint32_t *>(row + (long long)(sx + 1) * 4));
                        *reinterpret_cast<uint2 *>(dstp) = v;
                    }
                }
            }
        }
        cp_async_commit_group();
    };
    auto drain = [&]() {  // all threads; callers put barriers around it
        const int n = nAmb;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int code = ambList[e];
            const int oy = yFirst + (code >> 9), ox = blockIdx.x * (128 * kOut) + (code & 511);
            exact_px<false>(p, s, ox, oy, dimg + (long long)oy * p.dstRowStride + (long long)ox * 4);
        }
    };
    stage_row(yFirst, stage);
#pragma unroll 1
    for (int y0 = yFirst; y0 < yLast; y0++) {
        const int par = (y0 - yFirst) & 1;
        const uint8_t *buf = stage + par * STAGEB;
        cp_async_wait_group<0>();   // this thread's copies of row y0 (issued during the previous iteration) have landed
        // thread 0's view of the queue may miss pushes of the previous row that are still in flight (<= 512), and
        // this row can add 512 more: ask for a drain while 1024 slots are still free.  The flag alternates between
        // two slots so that the write for row y0+1 cannot overtake a slow thread's read for row y0.
        if (threadIdx.x == 0) flush[par] = nAmb > kAmbCap - 1024;
        __syncthreads();  // row y0 is staged and visible; everyone is done with row y0-1 (its buffer and its pushes)
        // Prefetch row y0+1 into the buffer row y0-1 used — only now, after the barrier: issued any earlier, a fast
        // warp would overwrite pixels a slow warp is still reading (found by running under compute-sanitizer, whose
        // timing exposed the race as a parity failure).  [Three buffers with a prefetch distance of two rows were
        // measured no faster: 1.02 vs 1.00 ms per 8 images.]
        if (y0 + 1 < yLast) stage_row(y0 + 1, stage + (par ^ 1) * STAGEB);
        if (flush[par]) {      // block-uniform, rare
            drain();
            __syncthreads();
            if (threadIdx.x == 0) nAmb = 0;
            __syncthreads();
        }
        if (x0 < p.outW) {
            uint32_t outv[kOut];
            bool ambv[kOut];
            const int rowCode = (y0 - yFirst) << 9;
            if (x0 >= q.dLo && x0 + kOut <= q.dHi) {
                uint32_t raw[NIN];
                // window = staging pixels 16*t .. 16*t + NIN - 1: chunks t, t+1, t+2
                const uint8_t *wbase = buf + threadIdx.x * CHB;
#pragma unroll
                for (int v4 = 0; v4 < (NIN + 3) / 4; v4++) {
                    const int u = v4 * 4;
                    const uint4 t4 = *reinterpret_cast<const uint4 *>(wbase + (u >> 4) * CHB + (u & 15) * 4);
                    raw[u] = t4.x;
                    if (u + 1 < NIN) raw[u + 1] = t4.y;
                    if (u + 2 < NIN) raw[u + 2] = t4.z;
                    if (u + 3 < NIN) raw[u + 3] = t4.w;
                }
                int_ratio_window<R, T, 0, NIN>(raw, q, outv, ambv);
            } else {  // edge outputs (clipped / renormalised taps): straight to the exact queue
#pragma unroll
                for (int j = 0; j < kOut; j++) { ambv[j] = true; outv[j] = 0u; }
            }
#pragma unroll
            for (int j = 0; j < kOut; j++) {
                if (x0 + j < p.outW) {
                    if (ambv[j]) ambList[atomicAdd(&nAmb, 1)] = (unsigned short)(rowCode + threadIdx.x * kOut + j);
                    else *reinterpret_cast<uint32_t *>(dimg + (long long)y0 * p.dstRowStride + (long long)(x0 + j) * 4) = outv[j];
                }
            }
        }
    }
    __syncthreads();
    drain();
}

// ------------------------------------------------------------------------------------------------
// Horizontal pass, warp-autonomous (round 2).  The block-wide kernel above spends 16 % of its stall samples on its one
// barrier per row and stages half of its blocks (the first and the last of a 1920-output row) through a branchy
// per-copy path (profiles/r2s2_lanczos).  Here a one-warp block owns 128 adjacent outputs (kOut per lane) and walks
// kWRows rows; the source span of a row (128*R + window pixels, 16-byte aligned: the window of a lane starts LEAD
// pixels into its first quad) is prefetched kLzStages-1 rows ahead into the warp's own ring with 16-byte cp.async
// (quads that cross a row end: zero-filled by hand), one __syncwarp per row, no barrier.  All kOut results of a lane
// go out as ONE 128-bit store; ambiguous / edge outputs are queued per warp and overwritten by the exact FP64 path
// 32 at a time.
// ------------------------------------------------------------------------------------------------
#ifndef FB_LZ_WROWS
#define FB_LZ_WROWS 16
#endif
#ifndef FB_LZ_STAGES
#define FB_LZ_STAGES 3
#endif
#ifndef FB_LZ_MINB
#define FB_LZ_MINB 24
#endif
#ifndef FB_LZ_KO_DEFAULT
#define FB_LZ_KO_DEFAULT 4
#endif
#ifndef FB_LZ_MINB8
#define FB_LZ_MINB8 16
#endif
constexpr int kWRows = FB_LZ_WROWS;
constexpr int kLzStages = FB_LZ_STAGES;
constexpr int kLzQ = 32 * kOut + 32;   // a row adds at most 32*kOut entries to fewer than 32 leftovers

__device__ __forceinline__ void cp_async16_lz(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// Exact path for `take` queued outputs (one per lane).  Not inlined: one copy of the FP64 sequence in the kernel
// instead of one per call site keeps the hot loop's code small (the first version stalled 0.64 cycles per issue on
// instruction fetch, profiles/r2s2b_lanczos).
// Interior destination of an integer ratio: taps R*d + off + k, k < T, with the shared binary64 weights q.wd[] — the same
// operations in the same order as exact_px, without the CSR index / weight loads.
template <bool VERTICAL, int R, int T>
__device__ __forceinline__ void exact_px_interior(const IntRatioParams &q, const uint8_t *s, int x, int y, uint8_t *dpx) {
    const ResizeParams &p = q.base;
    const int s0 = R * (VERTICAL ? y : x) + q.off;
    const uint8_t *base = VERTICAL ? s + (long long)s0 * p.srcRowStride + (long long)x * 4 : s + (long long)y * p.srcRowStride + (long long)s0 * 4;
    const long long step = VERTICAL ? (long long)p.srcRowStride : 4;
    double r2 = 0.0, g2 = 0.0, b2 = 0.0, a2 = 0.0;
#pragma unroll
    for (int tb = 0; tb < T; tb += 8) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (tb + k < T) v[k] = __ldg(reinterpret_cast<const uint32_t *>(base + (tb + k) * step));
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (tb + k < T) {
                const double aw = __dmul_rn(byte_to_double(v[k] >> 24), q.wd[tb + k]);
                r2 = __dadd_rn(r2, __dmul_rn(byte_to_double(v[k] & 0xFF), aw));
                g2 = __dadd_rn(g2, __dmul_rn(byte_to_double((v[k] >> 8) & 0xFF), aw));
                b2 = __dadd_rn(b2, __dmul_rn(byte_to_double((v[k] >> 16) & 0xFF), aw));
                a2 = __dadd_rn(a2, aw);
            }
        }
    }
    finish_px(r2, g2, b2, a2, dpx);
}

// Exact path for `take` queued outputs (one per lane).  Not inlined: one copy of the FP64 sequence in the kernel
// instead of one per call site keeps the hot loop's code small (the first version stalled 0.64 cycles per issue on
// instruction fetch, profiles/r2s2b_lanczos).  Interior destinations (the ambiguous ones) take the table-free form,
// edge destinations (clipped, renormalised taps) the CSR rows.
template <bool VERTICAL, int R, int T>
__device__ __noinline__ void lz_exact_queue(const IntRatioParams &q, const uint8_t *s, uint8_t *dimg, const uint32_t *queue, int take) {
    const ResizeParams &p = q.base;
    const int lane = threadIdx.x & 31;
    if (lane < take) {
        const uint32_t code = queue[lane];
        const int ox = (int)(code & 0xFFFFu), oy = (int)(code >> 16);
        const int d = VERTICAL ? oy : ox;
        uint8_t *dpx = dimg + (long long)oy * p.dstRowStride + (long long)ox * 4;
        if (q.wdExact && d >= q.dLo && d < q.dHi) exact_px_interior<VERTICAL, R, T>(q, s, ox, oy, dpx);
        else exact_px<VERTICAL>(p, s, ox, oy, dpx);
    }
}

template <int R, int T, int LEAD, int KO, int STAGES, int MINB>
__global__ void __launch_bounds__(32, MINB) resize_h_int_ratio_warp_kernel(const __grid_constant__ IntRatioParams q) {
    constexpr int NIN = T + (KO - 1) * R;                // window of one lane
    constexpr int NRAW = (NIN + LEAD + 3) / 4 * 4;        // ... read as whole quads
    constexpr int STEP = R * KO;                          // source pixels between the windows of adjacent lanes = one chunk
    static_assert(STEP % 4 == 0 && 32 % (STEP / 4) == 0 && KO % 4 == 0, "chunks are whole quads, 32 quads are whole chunks");
    constexpr int QPC = STEP / 4;                         // 16-byte quads per chunk
    constexpr int SPAN = STEP * 31 + NRAW;                // staged pixels per row
    constexpr int QUADS = SPAN / 4;
    constexpr int NK = (QUADS + 31) / 32;                 // 16-byte copies per lane and row
    constexpr int CHUNKS = (SPAN + STEP - 1) / STEP;
    constexpr int CHB = STEP * 4 + 16;                    // chunk + 16 B pad: conflict-free LDS.128 at 1 chunk / lane
    constexpr int STAGEB = CHUNKS * CHB;
    const ResizeParams &p = q.base;
    __shared__ __align__(16) uint8_t stage[STAGES][STAGEB];
    __shared__ uint32_t ambQ[32 * KO + 32];   // a row adds at most 32*KO entries to fewer than 32 leftovers
    const int lane = threadIdx.x;
    const int img = blockIdx.z;
    const int xw = blockIdx.x * (32 * KO);              // first output of the warp
    const int x0 = xw + lane * KO;
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;
    uint8_t *dimg = p.dst + (long long)img * p.dstImgStride;
    const int yFirst = blockIdx.y * kWRows;
    const int yLast = min(yFirst + kWRows, p.outH);       // exclusive
    const int sBase = R * xw + q.off - LEAD;              // staging pixel u <-> source pixel sBase + u; multiple of 4
    const bool al16 = ((((uintptr_t)s + (long long)sBase * 4) | (uintptr_t)p.srcRowStride) & 15) == 0;
    const bool interiorSpan = al16 && sBase >= 0 && sBase + SPAN <= p.srcW;
    const bool dvec = ((((uintptr_t)p.dst | (uintptr_t)p.dstImgStride | (uintptr_t)p.dstRowStride) & 15) == 0) && x0 + KO <= p.outW;
    const uint32_t myStage = (uint32_t)__cvta_generic_to_shared(&stage[0][0]) + (lane / QPC) * CHB + (lane % QPC) * 16;
    const uint8_t *myRow = s + (long long)sBase * 4 + lane * 16;   // this lane's first quad of row 0
    auto stage_row = [&](int y, int slot) {
        if (y < yLast) {
            const uint8_t *g = myRow + (long long)y * p.srcRowStride;
            const uint32_t d = myStage + slot * STAGEB;
            if (interiorSpan) {   // warp-uniform: NK copies at immediate offsets (quad v = lane + 32k: chunk += 8k)
#pragma unroll
                for (int k = 0; k < NK; k++)
                    if (k < QUADS / 32 || lane < QUADS - 32 * k)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + k * (32 / QPC) * CHB), "l"(g + k * 512) : "memory");
            } else {              // first / last warp of a row or an unaligned source: zero outside the row
#pragma unroll 1
                for (int k = 0; k < NK; k++) {
                    const int v = lane + 32 * k;
                    if (v < QUADS) {
                        const int sx = sBase + 4 * v;
                        if (al16 && sx >= 0 && sx + 4 <= p.srcW) {
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + k * (32 / QPC) * CHB), "l"(g + k * 512) : "memory");
                        } else {
                            uint32_t t[4];
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                t[i] = (sx + i >= 0 && sx + i < p.srcW) ? __ldg(reinterpret_cast<const uint32_t *>(g + k * 512 + i * 4)) : 0u;
                            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(d + k * (32 / QPC) * CHB), "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]) : "memory");
                        }
                    }
                }
            }
        }
        cp_async_commit_group();
    };
    int nq = 0;   // queue length (warp-uniform, in a register: pushes are ballot-compacted, no atomics)
#pragma unroll
    for (int k = 0; k < STAGES - 1; k++) stage_row(yFirst + k, k);
    int slot = 0;
#pragma unroll 1
    for (int y0 = yFirst; y0 < yLast; y0++) {
        // the slot row y0-1 used is free: every lane passed the __syncwarp at the end of that row
        stage_row(y0 + STAGES - 1, slot == 0 ? STAGES - 1 : slot - 1);
        cp_async_wait_group<STAGES - 1>();   // this lane's copies of row y0 have landed
        __syncwarp();                           // ... and everybody else's
        uint32_t outv[KO];
        bool ambv[KO];
        if (x0 >= q.dLo && x0 + KO <= q.dHi) {
            uint32_t raw[NRAW];
            const uint8_t *wbase = &stage[slot][0] + lane * CHB;   // staging pixels STEP*lane .. STEP*lane + NRAW - 1
#pragma unroll
            for (int v4 = 0; v4 < NRAW / 4; v4++) {
                const int u = v4 * 4;
                const uint4 t4 = *reinterpret_cast<const uint4 *>(wbase + (u / STEP) * CHB + (u % STEP) * 4);
                raw[u] = t4.x; raw[u + 1] = t4.y; raw[u + 2] = t4.z; raw[u + 3] = t4.w;
            }
            int_ratio_window<R, T, LEAD, NRAW, KO>(raw, q, outv, ambv);
        } else {  // edge outputs (clipped / renormalised taps; or beyond the row): straight to the exact queue
#pragma unroll
            for (int j = 0; j < KO; j++) { ambv[j] = true; outv[j] = 0u; }
        }
        uint8_t *drow = dimg + (long long)y0 * p.dstRowStride + (long long)x0 * 4;
        if (dvec) {   // queued ones are overwritten below
#pragma unroll
            for (int v = 0; v < KO / 4; v++)
                *reinterpret_cast<uint4 *>(drow + 16 * v) = make_uint4(outv[4 * v], outv[4 * v + 1], outv[4 * v + 2], outv[4 * v + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < KO; j++)
                if (x0 + j < p.outW) *reinterpret_cast<uint32_t *>(drow + j * 4) = outv[j];
        }
#pragma unroll
        for (int j = 0; j < KO; j++) {
            const bool push = ambv[j] && x0 + j < p.outW;
            const uint32_t b = __ballot_sync(0xffffffffu, push);
            if (b) {   // warp-uniform
                if (push) ambQ[nq + __popc(b & ((1u << lane) - 1u))] = ((uint32_t)y0 << 16) | (uint32_t)(x0 + j);
                nq += __popc(b);
            }
        }
        __syncwarp();   // stores, pushes and reads of stage[slot] are done before the drain / the next restage
        while (nq >= 32) {
            nq -= 32;
            lz_exact_queue<false, R, T>(q, s, dimg, ambQ + nq, 32);
        }
        slot = slot == STAGES - 1 ? 0 : slot + 1;
    }
    if (nq > 0) lz_exact_queue<false, R, T>(q, s, dimg, ambQ, nq);
}

// ------------------------------------------------------------------------------------------------
// Vertical pass, warp-autonomous (round 2).  resize_v_int_ratio_kernel above issues its 36-row window as 36 global loads
// with 64-bit address arithmetic each (144 of the 644 instructions of its main block) and synchronises the block twice
// for the exact queue (20 % of its stall samples at the two barriers, profiles/r2s2b_lanczos).  Here a one-warp block
// owns 32 adjacent columns and walks kVSteps steps of kOut output rows; the window of the NEXT step is copied into the
// second of two shared-memory buffers with 16-byte cp.async (8 lanes per row, 9 copies per lane) while the current
// one is evaluated from LDS at immediate offsets; the exact queue is per warp and ballot-compacted.
// Needs 16-byte-aligned rows and outW % 32 == 0 (else the kernel above runs).
// ------------------------------------------------------------------------------------------------
#ifndef FB_LZ_VSTEPS
#define FB_LZ_VSTEPS 12
#endif
constexpr int kVSteps = FB_LZ_VSTEPS;

template <int R, int T>
__global__ void __launch_bounds__(32, 20) resize_v_int_ratio_warp_kernel(const __grid_constant__ IntRatioParams q) {
    constexpr int NIN = T + (kOut - 1) * R;               // source rows of one step's window
    constexpr int NK = (NIN * 8 + 31) / 32;               // 16-byte copies per lane and step
    const ResizeParams &p = q.base;
    __shared__ __align__(16) uint32_t win[2][NIN][32];
    __shared__ uint32_t ambQ[kLzQ];
    const int lane = threadIdx.x;
    const int img = blockIdx.z;
    const int xw = blockIdx.x * 32, x0 = xw + lane;
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;
    uint8_t *dimg = p.dst + (long long)img * p.dstImgStride;
    const int yFirst = blockIdx.y * (kVSteps * kOut);
    const int yLast = min(yFirst + kVSteps * kOut, p.outH);   // exclusive
    const int srcH = p.srcW;                                   // launch_pass<true> passes the source HEIGHT as srcSize
    const uint8_t *colBase = s + (long long)xw * 4 + (lane & 7) * 16;
    const uint32_t myWin = (uint32_t)__cvta_generic_to_shared(&win[0][0][0]) + (lane >> 3) * 128 + (lane & 7) * 16;
    auto stage_step = [&](int y0, int buf) {   // window of outputs y0 .. y0+kOut-1: source rows R*y0 + off + i, clamped into the image
        if (y0 < yLast) {
            const int s0 = R * y0 + q.off;
#pragma unroll
            for (int k = 0; k < NK; k++) {
                const int r = (lane >> 3) + 4 * k;
                if (k < NIN / 4 || r < NIN) {
                    const int sy = min(max(s0 + r, 0), srcH - 1);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(myWin + buf * (NIN * 128) + k * 512),
                                 "l"(colBase + (long long)sy * p.srcRowStride) : "memory");
                }
            }
        }
        cp_async_commit_group();
    };
    int nq = 0;
    stage_step(yFirst, 0);
    int buf = 0;
#pragma unroll 1
    for (int y0 = yFirst; y0 < yLast; y0 += kOut, buf ^= 1) {
        stage_step(y0 + kOut, buf ^ 1);   // its previous reader (step y0 - kOut) is behind the __syncwarp at the end of that step
        cp_async_wait_group<1>();
        __syncwarp();
        uint32_t outv[kOut];
        bool ambv[kOut];
        if (y0 >= q.dLo && y0 + kOut <= q.dHi) {
            uint32_t raw[NIN];
#pragma unroll
            for (int i = 0; i < NIN; i++) raw[i] = win[buf][i][lane];
            int_ratio_window<R, T, 0, NIN>(raw, q, outv, ambv);
        } else {  // edge rows (clipped / renormalised taps): straight to the exact queue
#pragma unroll
            for (int j = 0; j < kOut; j++) { ambv[j] = true; outv[j] = 0u; }
        }
#pragma unroll
        for (int j = 0; j < kOut; j++) {
            const bool live = y0 + j < p.outH;
            if (live) *reinterpret_cast<uint32_t *>(dimg + (long long)(y0 + j) * p.dstRowStride + (long long)x0 * 4) = outv[j];   // queued ones are overwritten below
            const bool push = ambv[j] && live;
            const uint32_t b = __ballot_sync(0xffffffffu, push);
            if (b) {   // warp-uniform
                if (push) ambQ[nq + __popc(b & ((1u << lane) - 1u))] = ((uint32_t)(y0 + j) << 16) | (uint32_t)x0;
                nq += __popc(b);
            }
        }
        __syncwarp();
        while (nq >= 32) {
            nq -= 32;
            lz_exact_queue<true, R, T>(q, s, dimg, ambQ + nq, 32);
        }
    }
    if (nq > 0) lz_exact_queue<true, R, T>(q, s, dimg, ambQ, nq);
}

template <bool VERTICAL>
int launch_pass(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, uint8_t *dst,
                long long dstImgStride, int dstRowStride, int outW, int outH, int n, const int *start,
                const int *index, const double *weight, const float *weight32, int maxTaps, double wabs,
                const int *first = nullptr, const float *wpadT = nullptr, int groups = 0, int srcSize = 0,
                const IntRatioInfo *ir = nullptr) {
    if (n <= 0 || outW <= 0 || outH <= 0) return FB_OK;
    ResizeParams p;
    p.src = src; p.dst = dst;
    p.srcImgStride = srcImgStride; p.dstImgStride = dstImgStride;
    p.srcRowStride = srcRowStride; p.dstRowStride = dstRowStride;
    p.outW = outW; p.outH = outH;
    p.start = start; p.index = index; p.weight = weight; p.weight32 = weight32;
    // FMA k rounds a partial sum bounded by 255*255*P_k (colour) or 255*P_k (alpha), P_k = the partial sums of |w|; aw =
    // alpha*w and the FP32 weights add 2 more relative roundings on every term: wabs = max over destinations of
    // 2*sum|w| + sum_k P_k (api.cu table_stats).  15 % margin.
    const double u = 5.9604644775390625e-08;  // 2^-24
    (void)maxTaps;
    p.Er = (float)(u * 65025.0 * wabs * 1.15);
    p.Ea = (float)(u * 255.0 * wabs * 1.15);
    dim3 grid((outW + 255) / 256, outH, n);
    p.first = first; p.wpadT = wpadT; p.groups = groups; p.srcW = srcSize;
    p.vecOK = (((uintptr_t)src | (uintptr_t)srcImgStride | (uintptr_t)srcRowStride) & 15) == 0;
    if (weight32 != nullptr && p.Ea < 0.2f && getenv("FB_RESIZE_GENERIC") == nullptr && ir && ir->ratio >= 2 &&
        (VERTICAL || wpadT != nullptr) && getenv("FB_RESIZE_NO_INTRATIO") == nullptr) {
        IntRatioParams q;
        q.base = p; q.off = ir->off; q.dLo = ir->dLo; q.dHi = ir->dHi; q.wsum = ir->wsum;
        for (int i = 0; i < 28; i++) { q.w[i] = i < ir->taps ? ir->w[i] : 0.f; q.wn[i] = i < ir->taps ? ir->wn[i] : 0.f; }
        q.Eo = ir->Eo;
        q.wdExact = (ir->wdExact && getenv("FB_LZ_NO_WD") == nullptr) ? 1 : 0;
        for (int i = 0; i < 24; i++) q.wd[i] = i < ir->taps ? ir->wd[i] : 0.0;
        q.opaqueA = getenv("FB_LZ_NO_OPAQUE") == nullptr ? ir->opaqueA : 0u;
        dim3 g2 = VERTICAL ? dim3((outW + 127) / 128, (outH + kOut - 1) / kOut, n)
                           : dim3(((outW + kOut - 1) / kOut + 127) / 128, (outH + kRowsPerBlock - 1) / kRowsPerBlock, n);
        bool launched = true;
        static const bool oldH = getenv("FB_LZ_OLD") != nullptr;   // round-1 block-wide kernel, kept for A/B runs
        static const int hko = [] { const char *e = getenv("FB_LZ_KO"); return (e && e[0] == '4') ? 4 : (e && e[0] == '8') ? 8 : FB_LZ_KO_DEFAULT; }();
        if (!VERTICAL && ir->ratio == 4 && ir->taps == 24 && ((ir->off % 4) + 4) % 4 == 2 && !oldH && outW < 65536 && outH < 65536) {
            if (hko == 8) {   // eight outputs per lane: 256 outputs per warp and row
                dim3 gw((outW + 32 * 8 - 1) / (32 * 8), (outH + kWRows - 1) / kWRows, n);
                resize_h_int_ratio_warp_kernel<4, 24, 2, 8, 2, FB_LZ_MINB8><<<gw, 32, 0, s>>>(q);
            } else {
                dim3 gw((outW + 32 * 4 - 1) / (32 * 4), (outH + kWRows - 1) / kWRows, n);
                resize_h_int_ratio_warp_kernel<4, 24, 2, 4, kLzStages, FB_LZ_MINB><<<gw, 32, 0, s>>>(q);
            }
        } else if (!VERTICAL && ir->ratio == 4 && ir->taps == 24 && (ir->off & 1) == 0) resize_h_int_ratio_kernel<4, 24><<<g2, 128, 0, s>>>(q);
        else if (VERTICAL && ir->ratio == 4 && ir->taps == 24 && !oldH && outW % 32 == 0 && p.vecOK && outW < 65536 && outH < 65536) {
            dim3 gw(outW / 32, (outH + kVSteps * kOut - 1) / (kVSteps * kOut), n);
            resize_v_int_ratio_warp_kernel<4, 24><<<gw, 32, 0, s>>>(q);
        }
        else if (VERTICAL && ir->ratio == 4 && ir->taps == 24) resize_v_int_ratio_kernel<4, 24><<<g2, 128, 0, s>>>(q);
        else if (VERTICAL && ir->ratio == 2 && ir->taps == 12) resize_v_int_ratio_kernel<2, 12><<<g2, 128, 0, s>>>(q);
        else if (VERTICAL && ir->ratio == 3 && ir->taps == 17) resize_v_int_ratio_kernel<3, 17><<<g2, 128, 0, s>>>(q);
        else launched = false;
        if (launched) {
            FB_LAUNCHED(1);
            FB_CUDA(cudaGetLastError());
            return FB_OK;
        }
    }
    if (weight32 != nullptr && p.Ea < 0.2f && getenv("FB_RESIZE_GENERIC") == nullptr) {
        if (!VERTICAL && wpadT != nullptr) resize_pass_fast_kernel<VERTICAL, true><<<grid, 256, 0, s>>>(p);
        else resize_pass_fast_kernel<VERTICAL, false><<<grid, 256, 0, s>>>(p);
    } else
        resize_pass_kernel<VERTICAL><<<grid, 256, 0, s>>>(p);
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

}  // namespace

int launch_resize_h(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int srcW,
                    int srcH, uint8_t *dst, long long dstImgStride, int dstRowStride, int dstW, int n,
                    const int *start_dev, const int *index_dev, const double *weight_dev,
                    const float *weight32_dev, int maxTaps, double wabs, const int *first_dev,
                    const float *wpadT_dev, int groups, const IntRatioInfo *ir) {
    return launch_pass<false>(s, src, srcImgStride, srcRowStride, dst, dstImgStride, dstRowStride, dstW, srcH, n,
                              start_dev, index_dev, weight_dev, weight32_dev, maxTaps, wabs, first_dev, wpadT_dev,
                              groups, srcW, ir);
}

int launch_resize_v(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int srcW,
                    int srcH, uint8_t *dst, long long dstImgStride, int dstRowStride, int dstH, int n,
                    const int *start_dev, const int *index_dev, const double *weight_dev,
                    const float *weight32_dev, int maxTaps, double wabs, const IntRatioInfo *ir) {
    return launch_pass<true>(s, src, srcImgStride, srcRowStride, dst, dstImgStride, dstRowStride, srcW, dstH, n,
                             start_dev, index_dev, weight_dev, weight32_dev, maxTaps, wabs, nullptr, nullptr, 0, srcH, ir);
}

}  // namespace fb
