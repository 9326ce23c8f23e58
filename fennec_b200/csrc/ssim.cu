// ssim.cu — K1: BT.601 luma + 8x8 Gaussian-windowed SSIM + reduction, one pass over HBM.
//
// Replaces toLuminance + windowedSSIM (ssim.go:73-166, 207-220) and pixelSSIM (ssim.go:169-204).
//
// Formulation (DESIGN.md "K1"): the reference's two-pass 64-tap window is restated as a separable
// one-pass filter of four planes per pixel — a', b', q = a'^2 + b'^2, p = a'b' — where a' = luma(a) - c
// is centred on a per-strip constant c (the luma of the strip's centre pixel) so that
// E[x^2] - mu^2 does not cancel catastrophically in FP32 (measured <= 2e-6 absolute on adversarial
// inputs, tests/test_parity_gpu.py::test_scores_match_golden and tools/quick_ssim.py: <= 3.3e-7; the contract is
// 1e-5).  The window is
// exp(-(x^2+y^2)/4.5) for x,y in [-4,3] (ssim.go:74-77,116-117,223-241): an outer product g(y)g(x).
//
// Mapping: one warp owns a strip of 32*CPL input columns and walks down RS(+7) rows.  Each lane owns
// CPL adjacent columns, keeps the last 8 rows of its 4 planes in registers (the vertical 8-tap pass
// costs no memory traffic), publishes the vertical sums to shared memory once per row (conflict-free
// float4 layout) and reads its 7 right-hand neighbours back for the horizontal 8-tap pass.  Pixels
// are loaded with 128-bit non-allocating loads, one row ahead.  Bytes -> float goes through
// dp2a (integer dot product with the BT.601 weights x1000) accumulating straight into the bit pattern
// of 2^23 + L, because I2F runs at 1/8 of the FMA rate on sm_100 (tools/microbench.cu).
// This kernel is FMA-pipe-bound, not HBM-bound: ~92 FP32 lane-ops per pixel, FMA pipe 79 % busy
// (profiles/r2_k1_ncu_summaries.txt).
#include "common.cuh"

#include <cuda.h>   // CUtensorMap + the cuTensorMapEncodeTiled prototype (resolved at run time through cudart, no libcuda link)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace fb {

namespace {

constexpr float kC1f = 6.5025f;   // ssim.go:15
constexpr float kC2f = 58.5225f;  // ssim.go:16
constexpr float kLumaScale = 0.001f;

struct SsimParams {
    const uint8_t *a;
    const uint8_t *b;
    long long imgStrideA, imgStrideB;
    int rowStrideA, rowStrideB;
    int w, h, n;
    int nsx, nsy, rs;
    int vecOK;          // base pointers and strides are 16-byte aligned
    double *partials;   // [n][nsy*nsx]
    float g[8];         // 1-D Gaussian, k = -4..3, normalised
};

// 299*R + 587*G + 114*B accumulated onto the bits of 2^23 → float(2^23 + L), L <= 255000 < 2^23.
__device__ __forceinline__ float luma_magic(uint32_t px) {
    const uint32_t wRG = 299u | (587u << 16);
    const uint32_t wB0 = 114u;
    uint32_t t = __dp2a_hi(wB0, px, 0x4B000000u);
    t = __dp2a_lo(wRG, px, t);
    return __uint_as_float(t);
}

// ---- per-lane async global→shared copies (LDGSTS) with commit/wait groups -------------------------
// A 1-D TMA bulk-copy + mbarrier ring was measured first (round 1; the kernel is no longer in the tree — its tiled
// successor ssim_strip_tma_kernel below is, with its numbers in profiles/r2_k1_variants.txt): it removed the
// long-scoreboard stall too, but the elected-lane issue path (R2UR/UBLKCP/expect_tx) cost ~60 extra
// instructions per row per warp — more than it saved for 512-byte rows.  cp.async needs 4.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const uint8_t *src) {
    static_assert(BYTES == 16 || BYTES == 12 || BYTES == 8 || BYTES == 4, "cp.async copies 4, 8 or 16 bytes");
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else if (BYTES == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
    else if (BYTES == 12) {   // 4-byte aligned only
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4), "l"(src + 4) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 8), "l"(src + 8) : "memory");
    } else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// CPL packed pixels of one row for this lane.  FAST: one aligned vector load (the warp-uniform common
// case); otherwise per-pixel guarded loads (last strip of images whose width is not a multiple of 4,
// or unaligned buffers).
template <int CPL, bool FAST>
__device__ __forceinline__ void load_px(const uint8_t *p, int nvalid, uint32_t (&o)[CPL]) {
    if (FAST) {
        if (CPL == 4) {
            uint4 v = ld_nc_u128(p);
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[CPL - 1] = v.w;
        } else {
            uint2 v = ld_nc_u64(p);
            o[0] = v.x; o[CPL - 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < CPL; i++) o[i] = (i < nvalid) ? ld_nc_u32(p + 4 * i) : 0u;
    }
}

template <int CPL>
struct StripCtx {
    const uint8_t *pa, *pb;
    int rowStrideA, rowStrideB;
    int nIn, nvalid, lane;
    float K, c;
    float g[8];
    bool valid[CPL];
    float4 *vb0, *vb1;
    // FAST path only: per-warp ring of kStages row buffers filled by cp.async.
    uint8_t *ring;        // [kStages][2 images][32*CPL*4 bytes]
    int pxoff[CPL];       // CPL = 3 (4-byte copies): byte offset of each of the lane's pixels, clamped into the row
};

#ifndef FB_SSIM_F64FORMULA
#define FB_SSIM_F64FORMULA 0
#endif
#ifndef FB_SSIM_MINB4
#define FB_SSIM_MINB4 2
#endif
#ifndef FB_SSIM_STAGES
#define FB_SSIM_STAGES 4
#endif
#ifndef FB_SSIM_MINB1
#define FB_SSIM_MINB1 8    // one-warp blocks, CPL = 4: resident blocks the register allocation aims at (10 -> <= 200 registers)
#endif
#ifndef FB_SSIM_MINB3
#define FB_SSIM_MINB3 12   // CPL = 3: one-warp blocks per SM the register allocation aims at (12 -> <= 168 registers)
#endif
constexpr int kStages = FB_SSIM_STAGES;  // rows in flight per warp (power of two: the stage is the ring slot & (kStages-1))

// The row walk of one strip segment; returns this lane's sum of ssim/4 over its valid outputs.
template <int CPL, bool FAST>
__device__ __forceinline__ double walk_strip(const StripCtx<CPL> &q) {
    const uint8_t *pa = q.pa, *pb = q.pb;
    const int nIn = q.nIn, lane = q.lane;
    const float c = q.c;
    const float2 s2 = make_float2(kLumaScale, kLumaScale);
    const float2 K2 = make_float2(q.K, q.K);
    float2 g2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) g2[j] = make_float2(q.g[j], q.g[j]);
    // formula constants: th = c*(mua'+mub') + (c^2 + C1/2)
    const float kTh = fmaf(c, c, 0.5f * kC1f);
    const float2 qpInit = make_float2(kC2f, 0.5f * kC2f);  // Q' = Q + C2, P' = P + C2/2

    float2 rab[8][CPL], rqp[8][CPL];
    constexpr uint32_t kRowBuf = 32 * CPL * 4;  // bytes per image per stage
    // FAST: each lane copies its own CPL pixels of both images into its slot of a shared-memory ring with
    // cp.async, kStages rows ahead, and reads them back after cp.async.wait_group — no cross-lane
    // traffic, so no barrier is needed.  Register prefetch could not give that distance: all LDGs share
    // scoreboard slots, so waiting for row r also waited for the loads of rows r+1, r+2 issued after it
    // (profiles/r1_ssim_strip_ncu_summaries.txt, section r1d: 23% of samples in long-scoreboard on the first consumer).
    // !FAST: two rows in flight in registers, ping-pong by row parity.
    uint32_t pfa[2][CPL], pfb[2][CPL];
    const uint32_t myRing = FAST ? smem_u32(q.ring) + lane * (CPL * 4) : 0u;
    const uint8_t *myRingP = q.ring + lane * (CPL * 4);
    if (FAST) {
#pragma unroll
        for (int row = 0; row < kStages; row++) {  // rows 0..3 exist (nIn >= 8)
            cp_async<CPL * 4>(myRing + (2 * row) * kRowBuf, pa);
            cp_async<CPL * 4>(myRing + (2 * row + 1) * kRowBuf, pb);
            cp_async_commit();
            pa += q.rowStrideA;
            pb += q.rowStrideB;
        }
    } else {
        load_px<CPL, false>(pa, q.nvalid, pfa[0]);
        load_px<CPL, false>(pb, q.nvalid, pfb[0]);
        pa += q.rowStrideA;
        pb += q.rowStrideB;
        load_px<CPL, false>(pa, q.nvalid, pfa[1]);  // row 1 always exists (nIn >= 8)
        load_px<CPL, false>(pb, q.nvalid, pfb[1]);
    }

    // FB_FETCH(S, R): make row R's pixels available in pfa/pfb[(S)&1].  Groups complete in row order,
    // and one group is committed per row (empty at the tail), so "all but the newest kStages-1" == row R.
#define FB_FETCH(S, R)                                                                          \
    if (FAST) {                                                                                 \
        cp_async_wait<kStages - 1>();                                                           \
        const uint8_t *rb_ = myRingP + (2 * ((S) & (kStages - 1))) * kRowBuf;                   \
        if (CPL == 4) {                                                                         \
            uint4 va_ = *reinterpret_cast<const uint4 *>(rb_);                                  \
            uint4 vb_ = *reinterpret_cast<const uint4 *>(rb_ + kRowBuf);                        \
            pfa[(S) & 1][0] = va_.x; pfa[(S) & 1][1] = va_.y; pfa[(S) & 1][2] = va_.z; pfa[(S) & 1][CPL - 1] = va_.w; \
            pfb[(S) & 1][0] = vb_.x; pfb[(S) & 1][1] = vb_.y; pfb[(S) & 1][2] = vb_.z; pfb[(S) & 1][CPL - 1] = vb_.w; \
        } else {                                                                                \
            uint2 va_ = *reinterpret_cast<const uint2 *>(rb_);                                  \
            uint2 vb_ = *reinterpret_cast<const uint2 *>(rb_ + kRowBuf);                        \
            pfa[(S) & 1][0] = va_.x; pfa[(S) & 1][CPL - 1] = va_.y;                             \
            pfb[(S) & 1][0] = vb_.x; pfb[(S) & 1][CPL - 1] = vb_.y;                             \
        }                                                                                       \
    }

    // FB_REFILL(S, R): after row R has been consumed, start fetching row R+kStages into the same slot.
#define FB_REFILL(S, R)                                                                         \
    if (FAST) {                                                                                 \
        if ((R) + kStages < nIn) {                                                              \
            cp_async<CPL * 4>(myRing + (2 * ((S) & (kStages - 1))) * kRowBuf, pa);              \
            cp_async<CPL * 4>(myRing + (2 * ((S) & (kStages - 1)) + 1) * kRowBuf, pb);          \
            pa += q.rowStrideA;                                                                 \
            pb += q.rowStrideB;                                                                 \
        }                                                                                       \
        cp_async_commit();                                                                      \
    } else {                                                                                    \
        if ((R) + 2 < nIn) {                                                                    \
            pa += q.rowStrideA;                                                                 \
            pb += q.rowStrideB;                                                                 \
            load_px<CPL, false>(pa, q.nvalid, pfa[(S) & 1]);                                    \
            load_px<CPL, false>(pb, q.nvalid, pfb[(S) & 1]);                                    \
        }                                                                                       \
    }

#define FB_PLANES(S, R)                                                                         \
    FB_FETCH(S, R)                                                                              \
    _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                           \
        float2 f = make_float2(luma_magic(pfa[(S) & 1][i]), luma_magic(pfb[(S) & 1][i]));       \
        float2 t = __ffma2_rn(f, s2, K2);                                                       \
        float2 sq = __fmul2_rn(t, t);                                                           \
        rab[S][i] = t;                                                                          \
        rqp[S][i] = make_float2(sq.x + sq.y, t.x * t.y);                                        \
    }                                                                                           \
    FB_REFILL(S, R)

    // warm-up: input rows 0..6 fill ring slots 0..6
#pragma unroll
    for (int r = 0; r < 7; r++) {
        FB_PLANES(r, r)
    }

    const float2 one_two = make_float2(1.f, 2.f), neg2 = make_float2(-1.f, -1.f);
    float fs[CPL];  // per-lane sums of ssim/4 (<= 128 rows * 0.25 each: FP32 is ample)
#pragma unroll
    for (int i = 0; i < CPL; i++) fs[i] = 0.f;
    constexpr int kVLanes = 36;               // 32 lanes + pad for the 7-column right halo (ceil(7/CPL) <= 4)
    constexpr int kVBuf = CPL * kVLanes;      // float4 per buffer
    float4 *myV = q.vb0 + lane + ((7 & 1) ? kVBuf : 0);  // first main-loop row is r = 7 (odd)

    // One iteration per input row r >= 7: its planes go to ring slot r&7 and output row r-7 is produced.
    // Only the slot-specific part (planes + vertical taps, whose register indices depend on r&7) is
    // replicated by the switch; the horizontal pass and the SSIM formula exist once.
#pragma unroll 1
    for (int r = 7; r < nIn; r++) {
        float4 it[CPL + 7];

#define FB_SSIM_STEP(S)                                                                         \
    case S: {                                                                                   \
        FB_PLANES(S, r)                                                                         \
        /* rows r-7..r live in slots (S+1+j)&7, j = 0..7 */                                     \
        _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                       \
            float2 vab = __fmul2_rn(rab[(S + 1) & 7][i], g2[0]);                                \
            float2 vqp = __fmul2_rn(rqp[(S + 1) & 7][i], g2[0]);                                \
            _Pragma("unroll") for (int j = 1; j < 8; j++) {                                     \
                vab = __ffma2_rn(rab[(S + 1 + j) & 7][i], g2[j], vab);                          \
                vqp = __ffma2_rn(rqp[(S + 1 + j) & 7][i], g2[j], vqp);                          \
            }                                                                                   \
            it[i] = make_float4(vab.x, vab.y, vqp.x, vqp.y);                                    \
        }                                                                                       \
    } break;

        switch (r & 7) {
            FB_SSIM_STEP(0) FB_SSIM_STEP(1) FB_SSIM_STEP(2) FB_SSIM_STEP(3)
            FB_SSIM_STEP(4) FB_SSIM_STEP(5) FB_SSIM_STEP(6) FB_SSIM_STEP(7)
        }
#undef FB_SSIM_STEP

        // V rows are padded to kVLanes lane slots so neighbour reads need no clamp: item k of this lane is
        // at a compile-time offset from its own slot (lanes 30/31 read the pad; their outputs are masked).
        float4 *vb = myV;
        myV = (r & 1) ? myV - kVBuf : myV + kVBuf;  // double buffer (row parity)
#pragma unroll
        for (int i = 0; i < CPL; i++) vb[i * kVLanes] = it[i];
        __syncwarp();
#pragma unroll
        for (int k = CPL; k < CPL + 7; k++) it[k] = vb[(k % CPL) * kVLanes + k / CPL];
        // horizontal 8-tap + SSIM
#pragma unroll
        for (int i = 0; i < CPL; i++) {
            float2 mab = __fmul2_rn(make_float2(it[i].x, it[i].y), g2[0]);
            float2 mqp = __ffma2_rn(make_float2(it[i].z, it[i].w), g2[0], qpInit);
#pragma unroll
            for (int t = 1; t < 8; t++) {
                mab = __ffma2_rn(make_float2(it[i + t].x, it[i + t].y), g2[t], mab);
                mqp = __ffma2_rn(make_float2(it[i + t].z, it[i + t].w), g2[t], mqp);
            }
            // mab = (mua', mub'); mqp = (E[a'^2+b'^2] + C2, E[a'b'] + C2/2).  Formula packed WITHIN the pixel
            // (pairs (m,nn), (A1h,B1), (A2h,B2), (num,den)): 10 FMA-pipe instructions instead of 13, no transposes.
            // (Packing across two pixels was measured slower: the pair transposes cost more MOVs than they save.)
#if FB_SSIM_F64FORMULA
            // Experiment: the 13-op formula on the otherwise idle FP64 pipe (the FP32 FMA pipe is the binding resource).
            const double da = (double)mab.x, db = (double)mab.y;
            const double dm = da * db, dnn = fma(da, da, db * db);
            const double dth = fma((double)c, da + db, (double)kTh);
            const double n1 = dth + dm, d1 = fma(2.0, dth, dnn);
            const double n2 = (double)mqp.y - dm, d2 = (double)mqp.x - dnn;
            float ssim4 = __fdividef((float)(n1 * n2), (float)(d1 * d2));
            fs[i] += q.valid[i] ? ssim4 : 0.f;
#else
            const float2 sq = __fmul2_rn(mab, mab);                       // (mua'^2, mub'^2)
            const float2 mn = make_float2(mab.x * mab.y, sq.x + sq.y);    // (m, nn)
            const float th = fmaf(c, mab.x + mab.y, kTh);
            const float2 AB1 = __ffma2_rn(make_float2(th, th), one_two, mn);              // ((2 mua mub + C1)/2, mua^2+mub^2+C1)
            const float2 AB2 = __ffma2_rn(mn, neg2, make_float2(mqp.y, mqp.x));           // ((2 sab + C2)/2, saa+sbb+C2)
            const float2 nd = __fmul2_rn(AB1, AB2);                       // (num/4, den)
            float ssim4 = __fdividef(nd.x, nd.y);                         // ssim / 4
            fs[i] += q.valid[i] ? ssim4 : 0.f;
#endif
        }
    }
#undef FB_PLANES
#undef FB_FETCH
#undef FB_REFILL
    double dsum = 0.0;
#pragma unroll
    for (int i = 0; i < CPL; i++) dsum += (double)fs[i];
    return dsum;
}

// ------------------------------------------------------------------------------------------------
// Software-pipelined row walk (aligned inputs, CPL = 4).  walk_strip above runs, per row, a strictly serial chain:
// pixel LDS -> planes -> vertical taps -> STS -> __syncwarp -> neighbour LDS -> horizontal taps -> formula, and with
// two warps per scheduler the two shared-memory round trips are exposed.  Here the work of row r+1 that does not
// depend on the exchange is slotted into the waits of row r:
//     1. STS V(r); __syncwarp; issue the 7 neighbour LDS of V(r)
//     2. planes(r+1) from pixels fetched one step earlier          <- covers the neighbour-LDS latency
//     3. horizontal taps + formula of row r
//     4. cp.async.wait + pixel LDS of row r+2                       <- covered by step 5
//     5. vertical taps of row r+1 -> V(r+1)
// Register peak is unchanged (planes need no long-lived temporaries), the arithmetic is identical.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double walk_strip_pipe(const StripCtx<4> &q) {
    constexpr int CPL = 4;
    const uint8_t *pa = q.pa, *pb = q.pb;
    const int nIn = q.nIn, lane = q.lane;
    const float c = q.c;
    const float2 s2 = make_float2(kLumaScale, kLumaScale);
    const float2 K2 = make_float2(q.K, q.K);
    float2 g2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) g2[j] = make_float2(q.g[j], q.g[j]);
    const float kTh = fmaf(c, c, 0.5f * kC1f);
    const float2 qpInit = make_float2(kC2f, 0.5f * kC2f);
    float2 rab[8][CPL], rqp[8][CPL];
    constexpr uint32_t kRowBuf = 32 * CPL * 4;
    const uint32_t myRing = smem_u32(q.ring) + lane * (CPL * 4);
    const uint8_t *myRingP = q.ring + lane * (CPL * 4);
#pragma unroll
    for (int row = 0; row < kStages; row++) {  // rows 0..3 exist (nIn >= 8)
        cp_async<16>(myRing + (2 * row) * kRowBuf, pa);
        cp_async<16>(myRing + (2 * row + 1) * kRowBuf, pb);
        cp_async_commit();
        pa += q.rowStrideA;
        pb += q.rowStrideB;
    }
    uint4 pxa, pxb;   // pixels of the next row to convert
#define PIPE_FETCH(R)                                                                           \
    {                                                                                           \
        cp_async_wait<kStages - 1>();                                                           \
        const uint8_t *rb_ = myRingP + (2 * ((R) & (kStages - 1))) * kRowBuf;                   \
        pxa = *reinterpret_cast<const uint4 *>(rb_);                                            \
        pxb = *reinterpret_cast<const uint4 *>(rb_ + kRowBuf);                                  \
    }
    // planes of the fetched row into ring slot S, then start the copy of row R + kStages into the freed stage
#define PIPE_PLANES(S, R)                                                                       \
    {                                                                                           \
        const uint32_t xa_[4] = {pxa.x, pxa.y, pxa.z, pxa.w}, xb_[4] = {pxb.x, pxb.y, pxb.z, pxb.w}; \
        _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                       \
            float2 f = make_float2(luma_magic(xa_[i]), luma_magic(xb_[i]));                     \
            float2 t = __ffma2_rn(f, s2, K2);                                                   \
            float2 sq = __fmul2_rn(t, t);                                                       \
            rab[S][i] = t;                                                                      \
            rqp[S][i] = make_float2(sq.x + sq.y, t.x * t.y);                                    \
        }                                                                                       \
        if ((R) + kStages < nIn) {                                                              \
            cp_async<16>(myRing + (2 * ((R) & (kStages - 1))) * kRowBuf, pa);                   \
            cp_async<16>(myRing + (2 * ((R) & (kStages - 1)) + 1) * kRowBuf, pb);               \
            pa += q.rowStrideA;                                                                 \
            pb += q.rowStrideB;                                                                 \
        }                                                                                       \
        cp_async_commit();                                                                      \
    }
    // vertical taps for the row whose planes sit in slot S: rows r-7..r live in slots (S+1+j)&7
#define PIPE_VTAPS(S)                                                                           \
    _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                           \
        float2 vab = __fmul2_rn(rab[(S + 1) & 7][i], g2[0]);                                    \
        float2 vqp = __fmul2_rn(rqp[(S + 1) & 7][i], g2[0]);                                    \
        _Pragma("unroll") for (int j = 1; j < 8; j++) {                                         \
            vab = __ffma2_rn(rab[(S + 1 + j) & 7][i], g2[j], vab);                              \
            vqp = __ffma2_rn(rqp[(S + 1 + j) & 7][i], g2[j], vqp);                              \
        }                                                                                       \
        it[i] = make_float4(vab.x, vab.y, vqp.x, vqp.y);                                        \
    }

    float4 it[CPL + 7];
    // warm-up: rows 0..7 into slots 0..7, V(7), pixels of row 8
#pragma unroll
    for (int r = 0; r < 8; r++) {
        PIPE_FETCH(r)
        PIPE_PLANES(r, r)
    }
    PIPE_VTAPS(7)
    if (8 < nIn) PIPE_FETCH(8)

    const float2 one_two = make_float2(1.f, 2.f), neg2 = make_float2(-1.f, -1.f);
    float fs[CPL];
#pragma unroll
    for (int i = 0; i < CPL; i++) fs[i] = 0.f;
    constexpr int kVLanes = 36;
    constexpr int kVBuf = CPL * kVLanes;
    float4 *myV = q.vb0 + lane + ((7 & 1) ? kVBuf : 0);

#pragma unroll 1
    for (int r = 7; r < nIn; r++) {
        // 1. exchange V(r)
        float4 *vb = myV;
        myV = (r & 1) ? myV - kVBuf : myV + kVBuf;
#pragma unroll
        for (int i = 0; i < CPL; i++) vb[i * kVLanes] = it[i];
        __syncwarp();
#pragma unroll
        for (int k = CPL; k < CPL + 7; k++) it[k] = vb[(k % CPL) * kVLanes + k / CPL];
        // 2. planes of row r+1 (pixels fetched in step 4 of the previous iteration)
        const int rn = r + 1;
        if (rn < nIn) {
            switch (rn & 7) {
                case 0: PIPE_PLANES(0, rn) break;
                case 1: PIPE_PLANES(1, rn) break;
                case 2: PIPE_PLANES(2, rn) break;
                case 3: PIPE_PLANES(3, rn) break;
                case 4: PIPE_PLANES(4, rn) break;
                case 5: PIPE_PLANES(5, rn) break;
                case 6: PIPE_PLANES(6, rn) break;
                default: PIPE_PLANES(7, rn) break;
            }
        }
        // 3. horizontal 8-tap + SSIM of row r
#pragma unroll
        for (int i = 0; i < CPL; i++) {
            float2 mab = __fmul2_rn(make_float2(it[i].x, it[i].y), g2[0]);
            float2 mqp = __ffma2_rn(make_float2(it[i].z, it[i].w), g2[0], qpInit);
#pragma unroll
            for (int t = 1; t < 8; t++) {
                mab = __ffma2_rn(make_float2(it[i + t].x, it[i + t].y), g2[t], mab);
                mqp = __ffma2_rn(make_float2(it[i + t].z, it[i + t].w), g2[t], mqp);
            }
            const float2 sq = __fmul2_rn(mab, mab);
            const float2 mn = make_float2(mab.x * mab.y, sq.x + sq.y);
            const float th = fmaf(c, mab.x + mab.y, kTh);
            const float2 AB1 = __ffma2_rn(make_float2(th, th), one_two, mn);
            const float2 AB2 = __ffma2_rn(mn, neg2, make_float2(mqp.y, mqp.x));
            const float2 nd = __fmul2_rn(AB1, AB2);
            float ssim4 = __fdividef(nd.x, nd.y);
            fs[i] += q.valid[i] ? ssim4 : 0.f;
        }
        // 4. pixels of row r+2
        if (r + 2 < nIn) PIPE_FETCH(r + 2)
        // 5. vertical taps of row r+1
        if (rn < nIn) {
            switch (rn & 7) {
                case 0: PIPE_VTAPS(0) break;
                case 1: PIPE_VTAPS(1) break;
                case 2: PIPE_VTAPS(2) break;
                case 3: PIPE_VTAPS(3) break;
                case 4: PIPE_VTAPS(4) break;
                case 5: PIPE_VTAPS(5) break;
                case 6: PIPE_VTAPS(6) break;
                default: PIPE_VTAPS(7) break;
            }
        }
    }
#undef PIPE_FETCH
#undef PIPE_PLANES
#undef PIPE_VTAPS
    double dsum = 0.0;
#pragma unroll
    for (int i = 0; i < CPL; i++) dsum += (double)fs[i];
    return dsum;
}


// ------------------------------------------------------------------------------------------------
// Two rows per iteration (aligned inputs, CPL = 4).  walk_strip pays, per row, a three-level branch tree for the ring
// slot, a loop branch and one shared-memory round trip with nothing else to issue (profiles/r1_ssim_strip_ncu_summaries.txt: 12 % of the warp
// samples sit on BRA / ISETP / BSYNC / DEPBAR, and "wait" on fixed latencies is the largest stall).  Here one iteration
// takes rows r and r+1: four slot cases instead of eight, half the branches per row, both vertical passes back to back
// (16 independent FFMA2 chains), ONE __syncwarp for two rows, and the second row's neighbour loads are in flight
// while the first row's horizontal taps issue.  Four V buffers (two iterations x two rows) keep the single-barrier
// protocol of walk_strip valid.  An odd row count runs the second half of the last iteration on stale ring
// data (finite: real pixel rows) with its outputs masked.
// ------------------------------------------------------------------------------------------------
#ifndef FB_SSIM_PXMASK
#define FB_SSIM_PXMASK 1   // 1: mask every pixel's term (one FSEL); 0: mask columns at the end and skip rows by a uniform branch — one instruction per pixel less, yet 1.686 against 1.656 ms per 64 4K pairs (ptxas schedules the two differently)
#endif
struct HConsts {
    float2 cc;       // (c, c): the centring constant of the strip
    float2 c1;       // (C1/2, C1)
    float2 pqInit;   // (C2/2, C2): folded into the filter accumulators of the (p, q) planes
};
__device__ __forceinline__ HConsts make_hconsts(float c) {
    HConsts k;
    k.cc = make_float2(c, c);
    k.c1 = make_float2(0.5f * kC1f, kC1f);
    k.pqInit = make_float2(0.5f * kC2f, kC2f);
    return k;
}

// Horizontal 8-tap pass over the V rows + the SSIM formula.  With A, B the filtered centred lumas (mu - c) and
// P' = E[a'b'] + C2/2, Q' = E[a'^2 + b'^2] + C2 from the (p, q) planes:
//   (2 mua mub + C1)/2 = (A+c)(B+c) + C1/2         mua^2 + mub^2 + C1 = (A+c)^2 + (B+c)^2 + C1     (no cancellation: plain means)
//   (2 sab + C2)/2     = P' - A B                  saa + sbb + C2     = Q' - A^2 - B^2             (centred: E[x^2] - mu^2 stays small)
// as FADD2, FFMA2, FFMA, FFMA2, FFMA, two FMULs, MUFU.RCP, one FMA into the lane's sum: 11 FMA-pipe lane-ops per pixel
// (round 1: 14 — separate squares, m, nn, th and a fix-up inside __fdividef).
template <int CPL>
__device__ __forceinline__ void hpass_formula(const float4 (&own)[CPL], const float4 *vb, const float2 (&g2)[8],
                                              const HConsts &k, float (&fs)[CPL], const bool (&valid)[CPL], bool rowOK) {
    constexpr int kVLanes = 36;
    float4 it[CPL + 7];
#pragma unroll
    for (int i = 0; i < CPL; i++) it[i] = own[i];
#pragma unroll
    for (int t = CPL; t < CPL + 7; t++) it[t] = vb[(t % CPL) * kVLanes + t / CPL];
#pragma unroll
    for (int i = 0; i < CPL; i++) {
        float2 mab = __fmul2_rn(make_float2(it[i].x, it[i].y), g2[0]);
        float2 mpq = __ffma2_rn(make_float2(it[i].z, it[i].w), g2[0], k.pqInit);
#pragma unroll
        for (int t = 1; t < 8; t++) {
            mab = __ffma2_rn(make_float2(it[i + t].x, it[i + t].y), g2[t], mab);
            mpq = __ffma2_rn(make_float2(it[i + t].z, it[i + t].w), g2[t], mpq);
        }
        const float2 abc = __fadd2_rn(mab, k.cc);                                            // (mua, mub)
        const float2 n1 = __ffma2_rn(abc, make_float2(abc.y, abc.y), k.c1);                  // (mua mub + C1/2, mub^2 + C1)
        const float den1 = fmaf(abc.x, abc.x, n1.y);
        const float2 n2 = __ffma2_rn(make_float2(-mab.x, -mab.y), make_float2(mab.y, mab.y), mpq);   // (P' - AB, Q' - B^2)
        const float den2 = fmaf(-mab.x, mab.x, n2.y);
        const float num = n1.x * n2.x, den = den1 * den2;                                    // ssim / 4 = num / den
        float rden;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rden) : "f"(den));   // den in [C1*C2, ~1e10]: no range fix-up needed (__fdividef adds an FSETP and two predicated FMULs)
#if FB_SSIM_PXMASK
        fs[i] = fmaf((valid[i] && rowOK) ? num : 0.f, rden, fs[i]);   // per-pixel mask (FSEL), as round 1
#else
        fs[i] = fmaf(num, rden, fs[i]);   // columns without a valid output accumulate finite values the caller drops at the end
#endif
    }
}

// CPL = 3 (96 columns, 88 outputs per strip) trades 2.3 % more column halo and 4-byte copies (three per image and row,
// each pixel's address clamped into the row, so any width and any 4-byte-aligned buffer is "fast") for a 96-register
// ring: ~160 registers, three warps per scheduler instead of two.
template <int CPL>
__device__ __forceinline__ double walk_strip2(const StripCtx<CPL> &q) {
    static_assert(CPL == 3 || CPL == 4, "walk_strip2: 3 or 4 columns per lane");
    const uint8_t *pa = q.pa, *pb = q.pb;
    const int nIn = q.nIn, lane = q.lane;
    const float2 s2 = make_float2(kLumaScale, kLumaScale);
    const float2 K2 = make_float2(q.K, q.K);
    float2 g2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) g2[j] = make_float2(q.g[j], q.g[j]);
    HConsts hk;
    hk = make_hconsts(q.c);
    float2 rab[8][CPL], rqp[8][CPL];
    constexpr uint32_t kRowBuf = 32 * CPL * 4;
    const uint32_t myRing = smem_u32(q.ring) + lane * (CPL * 4);
    const uint8_t *myRingP = q.ring + lane * (CPL * 4);
    auto fetch_row = [&](int stage) {   // this lane's CPL pixels of both images, current row pointers
        if (CPL == 4) {
            cp_async<16>(myRing + (2 * stage) * kRowBuf, pa);
            cp_async<16>(myRing + (2 * stage + 1) * kRowBuf, pb);
        } else {
#pragma unroll
            for (int i = 0; i < CPL; i++) {
                cp_async<4>(myRing + (2 * stage) * kRowBuf + 4 * i, pa + q.pxoff[i]);
                cp_async<4>(myRing + (2 * stage + 1) * kRowBuf + 4 * i, pb + q.pxoff[i]);
            }
        }
        pa += q.rowStrideA;
        pb += q.rowStrideB;
    };
#pragma unroll
    for (int row = 0; row < kStages; row++) {  // rows 0..3 exist (nIn >= 8)
        fetch_row(row);
        cp_async_commit();
    }
    // planes of row R (ring stage R & 3) into slot S, then refill the stage with row R + kStages
#define W2_PLANES(S, R)                                                                         \
    {                                                                                           \
        cp_async_wait<kStages - 1>();                                                           \
        const uint8_t *rb_ = myRingP + (2 * ((S) & (kStages - 1))) * kRowBuf;                   \
        uint32_t xa_[4], xb_[4];                                                                \
        if (CPL == 4) {                                                                         \
            const uint4 va_ = *reinterpret_cast<const uint4 *>(rb_);                            \
            const uint4 vb_ = *reinterpret_cast<const uint4 *>(rb_ + kRowBuf);                  \
            xa_[0] = va_.x; xa_[1] = va_.y; xa_[2] = va_.z; xa_[3] = va_.w;                     \
            xb_[0] = vb_.x; xb_[1] = vb_.y; xb_[2] = vb_.z; xb_[3] = vb_.w;                     \
        } else {                                                                                \
            _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                   \
                xa_[i] = *reinterpret_cast<const uint32_t *>(rb_ + 4 * i);                      \
                xb_[i] = *reinterpret_cast<const uint32_t *>(rb_ + kRowBuf + 4 * i);            \
            }                                                                                   \
        }                                                                                       \
        _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                       \
            float2 f = make_float2(luma_magic(xa_[i]), luma_magic(xb_[i]));                     \
            float2 t = __ffma2_rn(f, s2, K2);                                                   \
            rab[S][i] = t;                                                                      \
            rqp[S][i] = make_float2(t.x * t.y, fmaf(t.x, t.x, t.y * t.y));   /* (p, q) */       \
        }                                                                                       \
        if ((R) + kStages < nIn) fetch_row((S) & (kStages - 1));                                \
        cp_async_commit();                                                                      \
    }
#define W2_VTAPS(S, V)                                                                          \
    _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                           \
        float2 vab = __fmul2_rn(rab[(S + 1) & 7][i], g2[0]);                                    \
        float2 vqp = __fmul2_rn(rqp[(S + 1) & 7][i], g2[0]);                                    \
        _Pragma("unroll") for (int j = 1; j < 8; j++) {                                         \
            vab = __ffma2_rn(rab[(S + 1 + j) & 7][i], g2[j], vab);                              \
            vqp = __ffma2_rn(rqp[(S + 1 + j) & 7][i], g2[j], vqp);                              \
        }                                                                                       \
        V[i] = make_float4(vab.x, vab.y, vqp.x, vqp.y);                                         \
    }
#pragma unroll
    for (int r = 0; r < 7; r++) W2_PLANES(r, r)

    float fs[CPL];
#pragma unroll
    for (int i = 0; i < CPL; i++) fs[i] = 0.f;
    constexpr int kVLanes = 36, kVBuf = CPL * kVLanes;

    // [Software-pipelined lumas — readback, dp2a and centring of rows r+2, r+3 issued in the basic block of the horizontal pass
    // of rows r, r+1 (slot-independent, so outside the switch), planes + vertical taps first thing in the next iteration —
    // were measured at 1.733 against 1.651 ms per 64 4K pairs (218 registers); as with MODE 1 in round 1, ptxas' own
    // interleaving of the serial version is better.  profiles/r2_tuning_sweep.txt]
#pragma unroll 1
    for (int r = 7; r < nIn; r += 2) {
        float4 vA[CPL], vB[CPL];
#define W2_STEP(S0, S1)                                                                         \
    case S0: {                                                                                  \
        W2_PLANES(S0, r)                                                                        \
        W2_VTAPS(S0, vA)                                                                        \
        W2_PLANES(S1, r + 1)                                                                    \
        W2_VTAPS(S1, vB)                                                                        \
    } break;
        switch (r & 7) {
            W2_STEP(7, 0) W2_STEP(1, 2) W2_STEP(3, 4)
            default: { W2_PLANES(5, r) W2_VTAPS(5, vA) W2_PLANES(6, r + 1) W2_VTAPS(6, vB) } break;
        }
#undef W2_STEP
        float4 *vbA = q.vb0 + ((((r - 7) >> 1) & 1) ? 2 * kVBuf : 0) + lane;
        float4 *vbB = vbA + kVBuf;
#pragma unroll
        for (int i = 0; i < CPL; i++) { vbA[i * kVLanes] = vA[i]; vbB[i * kVLanes] = vB[i]; }
        __syncwarp();
        hpass_formula<CPL>(vA, vbA, g2, hk, fs, q.valid, true);
        if (FB_SSIM_PXMASK || r + 1 < nIn) hpass_formula<CPL>(vB, vbB, g2, hk, fs, q.valid, r + 1 < nIn);   // warp-uniform: an odd row count ends on a half iteration
    }
#undef W2_PLANES
#undef W2_VTAPS
    double dsum = 0.0;
#pragma unroll
    for (int i = 0; i < CPL; i++) dsum += (FB_SSIM_PXMASK || q.valid[i]) ? (double)fs[i] : 0.0;
    return dsum;
}

// Four rows per iteration: two slot cases, a quarter of the branches per row.  V rows go to shared memory as soon as they
// are produced (own values are read back with the neighbours: 11 instead of 7 LDS.128 per lane and row), so the register
// peak stays that of one row; the NR buffers are single-buffered behind a second __syncwarp.
// NR = 2 with OWN-from-shared is the register-lean twin of walk_strip2.
template <int NR>
__device__ __forceinline__ double walk_stripN(const StripCtx<4> &q) {
    constexpr int CPL = 4;
    static_assert(NR == 2 || NR == 4, "slots of one iteration must not wrap inside a case");
    const uint8_t *pa = q.pa, *pb = q.pb;
    const int nIn = q.nIn, lane = q.lane;
    const float2 s2 = make_float2(kLumaScale, kLumaScale);
    const float2 K2 = make_float2(q.K, q.K);
    float2 g2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) g2[j] = make_float2(q.g[j], q.g[j]);
    HConsts hk;
    hk = make_hconsts(q.c);
    float2 rab[8][CPL], rqp[8][CPL];
    constexpr uint32_t kRowBuf = 32 * CPL * 4;
    const uint32_t myRing = smem_u32(q.ring) + lane * 16;
    const uint8_t *myRingP = q.ring + lane * 16;
#pragma unroll
    for (int row = 0; row < kStages; row++) {
        cp_async<16>(myRing + (2 * row) * kRowBuf, pa);
        cp_async<16>(myRing + (2 * row + 1) * kRowBuf, pb);
        cp_async_commit();
        pa += q.rowStrideA;
        pb += q.rowStrideB;
    }
#define WN_PLANES(S, R)                                                                         \
    {                                                                                           \
        cp_async_wait<kStages - 1>();                                                           \
        const uint8_t *rb_ = myRingP + (2 * ((S) & (kStages - 1))) * kRowBuf;                   \
        const uint4 va_ = *reinterpret_cast<const uint4 *>(rb_);                                \
        const uint4 vb_ = *reinterpret_cast<const uint4 *>(rb_ + kRowBuf);                      \
        const uint32_t xa_[4] = {va_.x, va_.y, va_.z, va_.w}, xb_[4] = {vb_.x, vb_.y, vb_.z, vb_.w}; \
        _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                       \
            float2 f = make_float2(luma_magic(xa_[i]), luma_magic(xb_[i]));                     \
            float2 t = __ffma2_rn(f, s2, K2);                                                   \
            rab[(S) & 7][i] = t;                                                                \
            rqp[(S) & 7][i] = make_float2(t.x * t.y, fmaf(t.x, t.x, t.y * t.y));   /* (p, q) */ \
        }                                                                                       \
        if ((R) + kStages < nIn) {                                                              \
            cp_async<16>(myRing + (2 * ((S) & (kStages - 1))) * kRowBuf, pa);                   \
            cp_async<16>(myRing + (2 * ((S) & (kStages - 1)) + 1) * kRowBuf, pb);               \
            pa += q.rowStrideA;                                                                 \
            pb += q.rowStrideB;                                                                 \
        }                                                                                       \
        cp_async_commit();                                                                      \
    }
    // vertical taps of the row in slot S, stored straight to V buffer J
#define WN_VTAPS_STS(S, J)                                                                      \
    _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                           \
        float2 vab = __fmul2_rn(rab[((S) + 1) & 7][i], g2[0]);                                  \
        float2 vqp = __fmul2_rn(rqp[((S) + 1) & 7][i], g2[0]);                                  \
        _Pragma("unroll") for (int j = 1; j < 8; j++) {                                         \
            vab = __ffma2_rn(rab[((S) + 1 + j) & 7][i], g2[j], vab);                            \
            vqp = __ffma2_rn(rqp[((S) + 1 + j) & 7][i], g2[j], vqp);                            \
        }                                                                                       \
        vb0[(J) * kVBuf + i * kVLanes] = make_float4(vab.x, vab.y, vqp.x, vqp.y);               \
    }
#pragma unroll
    for (int r = 0; r < 7; r++) WN_PLANES(r, r)

    float fs[CPL];
#pragma unroll
    for (int i = 0; i < CPL; i++) fs[i] = 0.f;
    constexpr int kVLanes = 36, kVBuf = CPL * kVLanes;
    float4 *vb0 = q.vb0 + lane;

#pragma unroll 1
    for (int r = 7; r < nIn; r += NR) {
        __syncwarp();   // every lane has finished reading the previous iteration's V rows
#define WN_ROW(S, J) WN_PLANES((S) + (J), r + (J)) WN_VTAPS_STS((S) + (J), J)
        if (NR == 4) {
            if ((r & 7) == 7) { WN_ROW(7, 0) WN_ROW(7, 1) WN_ROW(7, 2) WN_ROW(7, 3) }
            else              { WN_ROW(3, 0) WN_ROW(3, 1) WN_ROW(3, 2) WN_ROW(3, 3) }
        } else {
            switch (r & 7) {
                case 7: { WN_ROW(7, 0) WN_ROW(7, 1) } break;
                case 1: { WN_ROW(1, 0) WN_ROW(1, 1) } break;
                case 3: { WN_ROW(3, 0) WN_ROW(3, 1) } break;
                default: { WN_ROW(5, 0) WN_ROW(5, 1) } break;
            }
        }
#undef WN_ROW
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NR; j++) {
            if (r + j < nIn) {   // warp-uniform
                float4 own[CPL];
#pragma unroll
                for (int i = 0; i < CPL; i++) own[i] = vb0[j * kVBuf + i * kVLanes];
                hpass_formula(own, vb0 + j * kVBuf, g2, hk, fs, q.valid, true);
            }
        }
    }
#undef WN_PLANES
#undef WN_VTAPS_STS
    double dsum = 0.0;
#pragma unroll
    for (int i = 0; i < CPL; i++) dsum += (FB_SSIM_PXMASK || q.valid[i]) ? (double)fs[i] : 0.0;
    return dsum;
}

// ------------------------------------------------------------------------------------------------
// TMA variant of walk_strip2 (FB_SSIM_MODE=5): the pixel rows of a strip are staged by the Tensor Memory Accelerator —
// one elected lane issues two cp.async.bulk.tensor.3d copies (image a and image b, boxes of 128 pixels x 4 or 8 rows out
// of a {w, h, n} tensor map) per stage into a shared-memory ring of 16 rows and arms the stage's mbarrier with the
// byte count; the warp waits on the barrier's phase parity before the first row of a stage.  Compared with the
// per-lane cp.async ring this removes the per-row address arithmetic, LDGSTS, commit and wait_group of every lane
// (about 9 instructions per row) for about 4 per row in one lane, and the tensor map zero-fills everything right of /
// below the image, so ragged right edges need no re-aimed loads.  (Round 1 measured a 1-D bulk copy PER ROW slower
// than cp.async: the elected-lane issue path then cost ~60 instructions per 512-byte row; with 2-D boxes of four
// rows it is amortised.)  Everything after the pixel fetch is walk_strip2.
// ------------------------------------------------------------------------------------------------
// <rows per stage, stages in flight> (powers of two, rows <= 8): <4, 4> (FB_SSIM_MODE=5) and <8, 2> (FB_SSIM_MODE=6)

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TW_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TD_%=;\n"
        "bra TW_%=;\n"
        "TD_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

struct TmaCtx {
    const CUtensorMap *ta, *tb;
    int X0, Y0, img;          // strip origin (pixels / rows) and image index
    uint32_t ring, bars;      // shared-memory addresses: kTmaStages x kTmaStageBytes, kTmaStages mbarriers
    const uint8_t *ringP;
};

template <int kTmaRows, int kTmaStages>
__device__ __forceinline__ double walk_strip_tma(const StripCtx<4> &q, const TmaCtx &t) {
    constexpr int CPL = 4;
    constexpr uint32_t kTmaStageBytes = 2u * kTmaRows * 512u;   // both images
    const int nIn = q.nIn, lane = q.lane;
    const float2 s2 = make_float2(kLumaScale, kLumaScale);
    const float2 K2 = make_float2(q.K, q.K);
    float2 g2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) g2[j] = make_float2(q.g[j], q.g[j]);
    HConsts hk;
    hk = make_hconsts(q.c);
    float2 rab[8][CPL], rqp[8][CPL];
    // stage st <- rows [row0, row0 + kTmaRows) of both images (rows below the image arrive as zeros)
    auto issue = [&](int st, int row0) {
        const uint32_t bar = t.bars + 8u * st, dst = t.ring + (uint32_t)st * kTmaStageBytes;
        mbar_expect_tx(bar, kTmaStageBytes);
        tma_load_3d(dst, t.ta, t.X0, t.Y0 + row0, t.img, bar);
        tma_load_3d(dst + kTmaRows * 512u, t.tb, t.X0, t.Y0 + row0, t.img, bar);
    };
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < kTmaStages; st++)
            if (st * kTmaRows < nIn) issue(st, st * kTmaRows);
    }
    __syncwarp();
    const uint8_t *myP = t.ringP + lane * 16;
    // planes of row R (slot S = R & 7, so R's position inside its stage is known at compile time)
#define WT_PLANES(S, R)                                                                         \
    {                                                                                           \
        const int st_ = ((R) / kTmaRows) & (kTmaStages - 1);                                    \
        if (((S) & (kTmaRows - 1)) == 0) mbar_wait_parity(t.bars + 8u * st_, ((R) / (kTmaRows * kTmaStages)) & 1); \
        const uint8_t *rb_ = myP + (size_t)st_ * kTmaStageBytes + ((S) & (kTmaRows - 1)) * 512; \
        const uint4 va_ = *reinterpret_cast<const uint4 *>(rb_);                                \
        const uint4 vb_ = *reinterpret_cast<const uint4 *>(rb_ + kTmaRows * 512);               \
        const uint32_t xa_[4] = {va_.x, va_.y, va_.z, va_.w}, xb_[4] = {vb_.x, vb_.y, vb_.z, vb_.w}; \
        _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                       \
            float2 f = make_float2(luma_magic(xa_[i]), luma_magic(xb_[i]));                     \
            float2 tt = __ffma2_rn(f, s2, K2);                                                  \
            rab[S][i] = tt;                                                                     \
            rqp[S][i] = make_float2(tt.x * tt.y, fmaf(tt.x, tt.x, tt.y * tt.y));   /* (p, q) */ \
        }                                                                                       \
        if (((S) & (kTmaRows - 1)) == kTmaRows - 1) {   /* last row of the stage: every lane has read it; refill it */ \
            __syncwarp();                                                                       \
            const int next_ = (R) + 1 + kTmaRows * (kTmaStages - 1);                            \
            if (lane == 0 && next_ < nIn) issue(st_, next_);                                    \
        }                                                                                       \
    }
#define WT_VTAPS(S, V)                                                                          \
    _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                           \
        float2 vab = __fmul2_rn(rab[(S + 1) & 7][i], g2[0]);                                    \
        float2 vqp = __fmul2_rn(rqp[(S + 1) & 7][i], g2[0]);                                    \
        _Pragma("unroll") for (int j = 1; j < 8; j++) {                                         \
            vab = __ffma2_rn(rab[(S + 1 + j) & 7][i], g2[j], vab);                              \
            vqp = __ffma2_rn(rqp[(S + 1 + j) & 7][i], g2[j], vqp);                              \
        }                                                                                       \
        V[i] = make_float4(vab.x, vab.y, vqp.x, vqp.y);                                         \
    }
#pragma unroll
    for (int r = 0; r < 7; r++) WT_PLANES(r, r)

    float fs[CPL];
#pragma unroll
    for (int i = 0; i < CPL; i++) fs[i] = 0.f;
    constexpr int kVLanes = 36, kVBuf = CPL * kVLanes;

#pragma unroll 1
    for (int r = 7; r < nIn; r += 2) {
        float4 vA[CPL], vB[CPL];
        // A stage is only waited for when one of its rows exists (r + 1 < nIn), so the odd tail row reuses the planes of
        // an older row instead of waiting on a barrier nobody armed; its outputs are masked.
        const bool second = r + 1 < nIn;
#define WT_STEP(S0, S1)                                                                         \
    case S0: {                                                                                  \
        WT_PLANES(S0, r)                                                                        \
        WT_VTAPS(S0, vA)                                                                        \
        if (second) WT_PLANES(S1, r + 1)                                                        \
        WT_VTAPS(S1, vB)                                                                        \
    } break;
        switch (r & 7) {
            WT_STEP(7, 0) WT_STEP(1, 2) WT_STEP(3, 4)
            default: { WT_PLANES(5, r) WT_VTAPS(5, vA) if (second) WT_PLANES(6, r + 1) WT_VTAPS(6, vB) } break;
        }
#undef WT_STEP
        float4 *vbA = q.vb0 + ((((r - 7) >> 1) & 1) ? 2 * kVBuf : 0) + lane;
        float4 *vbB = vbA + kVBuf;
#pragma unroll
        for (int i = 0; i < CPL; i++) { vbA[i * kVLanes] = vA[i]; vbB[i * kVLanes] = vB[i]; }
        __syncwarp();
        hpass_formula(vA, vbA, g2, hk, fs, q.valid, true);
        if (second) hpass_formula(vB, vbB, g2, hk, fs, q.valid, true);   // warp-uniform
    }
#undef WT_PLANES
#undef WT_VTAPS
    double dsum = 0.0;
#pragma unroll
    for (int i = 0; i < CPL; i++) dsum += (FB_SSIM_PXMASK || q.valid[i]) ? (double)fs[i] : 0.0;
    return dsum;
}

// One warp per block = one strip segment; aligned inputs only (the host falls back to the cp.async kernel otherwise).
template <int kTmaRows, int kTmaStages>
__global__ void __launch_bounds__(32, 8) ssim_strip_tma_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb,
                                                               const SsimParams p) {
    constexpr int CPL = 4, OUTC = 120;
    constexpr uint32_t kTmaStageBytes = 2u * kTmaRows * 512u;
    __shared__ __align__(128) uint8_t ring[kTmaStages * kTmaStageBytes];
    __shared__ float4 vbuf[4][CPL * 36];
    __shared__ __align__(8) unsigned long long bars[kTmaStages];
    const int lane = threadIdx.x;
    const long long segsPerImg = (long long)p.nsx * p.nsy;
    const long long seg = blockIdx.x;
    const int img = (int)(seg / segsPerImg);
    const int rseg = (int)(seg - (long long)img * segsPerImg);
    const int sy = rseg / p.nsx, sx = rseg - sy * p.nsx;
    const int X0 = sx * OUTC, Y0 = sy * p.rs;
    const int nOut = min(p.rs, (p.h - 8) - Y0);
    const int xl = X0 + CPL * lane;
    if (lane < kTmaStages) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[lane])), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (lane < 4) {
#pragma unroll
        for (int i = 0; i < CPL; i++) {
#pragma unroll
            for (int v = 0; v < 4; v++) vbuf[v][i * 36 + 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncwarp();
    StripCtx<CPL> q;
    q.nIn = nOut + 7;
    q.lane = lane;
    q.nvalid = CPL;
    q.rowStrideA = p.rowStrideA;
    q.rowStrideB = p.rowStrideB;
    q.vb0 = &vbuf[0][0];
    q.vb1 = &vbuf[1][0];
    q.ring = ring;
    q.pa = q.pb = nullptr;
#pragma unroll
    for (int j = 0; j < 8; j++) q.g[j] = p.g[j];
#pragma unroll
    for (int i = 0; i < CPL; i++) q.valid[i] = (CPL * lane + i < OUTC) && (xl + i < p.w - 8);
    {   // centring constant: luma of image a at the strip's centre pixel (same rule as ssim_strip_kernel)
        const uint8_t *ia = p.a + (long long)img * p.imgStrideA;
        int cx = min(X0 + OUTC / 2, p.w - 1), cy = min(Y0 + q.nIn / 2, p.h - 1);
        float f = luma_magic(ld_nc_u32(ia + (long long)cy * p.rowStrideA + (long long)cx * 4));
        q.K = -(f * kLumaScale);
        q.c = -fmaf(8388608.0f, kLumaScale, q.K);
    }
    TmaCtx t;
    t.ta = &ta; t.tb = &tb;
    t.X0 = X0; t.Y0 = Y0; t.img = img;
    t.ring = smem_u32(ring);
    t.bars = smem_u32(bars);
    t.ringP = ring;
    double dsum = walk_strip_tma<kTmaRows, kTmaStages>(q, t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    if (lane == 0) p.partials[seg] = dsum * 4.0;
}

// Defaults from the round-2 sweep on 64 4K pairs (profiles/r2_k1_variants.txt): walk_strip / 4 warps per block 2.062 ms,
// walk_strip / 1 warp 1.882, walk_strip2 / 2 warps 1.956, walk_strip2 / 1 warp 1.769, walk_strip_pipe 2.272.
#ifndef FB_SSIM_MODE_DEFAULT
#define FB_SSIM_MODE_DEFAULT 2
#endif
#ifndef FB_SSIM_WPB_DEFAULT
#define FB_SSIM_WPB_DEFAULT 1
#endif

// MODE 0: walk_strip, 1: walk_strip_pipe, 2: walk_strip2, 3: walk_stripN<4>, 4: walk_stripN<2> (5: ssim_strip_tma_kernel).  WPB = warps (independent strips) per block: warps never
// synchronise with each other, so the block size only sets the granularity at which the SM's registers are handed out:
// 2 blocks of 4 warps at <= 255 registers, or one-warp blocks (ptxas settles at 200-208 registers with the 8-block hint; the
// register file is per scheduler, 16 K registers each, so anything from 171 to 255 registers means two warps per scheduler =
// 8 per SM — asking for 9 or 10 blocks makes ptxas cap at 168 registers, with spills).
template <int CPL, int WPB>
constexpr int ssim_min_blocks() { return CPL == 3 ? FB_SSIM_MINB3 / WPB : CPL != 4 ? 16 / WPB : (WPB == 4 ? FB_SSIM_MINB4 : WPB == 2 ? 4 : FB_SSIM_MINB1); }

#ifdef FB_SSIM_MAXNREG   // experiment: an explicit register cap instead of the resident-blocks hint
#define FB_SSIM_BOUNDS(CPL, WPB) __maxnreg__(FB_SSIM_MAXNREG)
#else
#define FB_SSIM_BOUNDS(CPL, WPB) __launch_bounds__(32 * WPB, (ssim_min_blocks<CPL, WPB>()))
#endif
template <int CPL, int MODE = 0, int WPB = 4>
__global__ void FB_SSIM_BOUNDS(CPL, WPB) ssim_strip_kernel(const SsimParams p) {
    constexpr int WARPS = WPB;
    constexpr bool PIPE = MODE == 1;
    constexpr int NVB = (MODE == 2 || MODE == 3) ? 4 : 2;   // MODE 4 (two rows, single-buffered) needs 2
    constexpr int INC = 32 * CPL;    // input columns per strip
    constexpr int OUTC = INC - 8;    // outputs per strip (multiple of 4 → 16-byte aligned strips)
    __shared__ float4 vbuf[WARPS][NVB][CPL * 36];   // padded V rows, see walk_strip
    __shared__ __align__(128) uint8_t pxring[WARPS][kStages][2][INC * 4];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long segsPerImg = (long long)p.nsx * p.nsy;
    const long long seg = (long long)blockIdx.x * WARPS + warp;
    if (seg >= segsPerImg * p.n) return;  // no block-wide barriers below
    const int img = (int)(seg / segsPerImg);
    const int rseg = (int)(seg - (long long)img * segsPerImg);
    const int sy = rseg / p.nsx, sx = rseg - sy * p.nsx;

    const int X0 = sx * OUTC, Y0 = sy * p.rs;
    const int nOut = min(p.rs, (p.h - 8) - Y0);  // output rows of this segment (>= 1)
    const int xl = X0 + CPL * lane;

    StripCtx<CPL> q;
    q.nIn = nOut + 7;
    q.lane = lane;
    q.nvalid = max(0, min(CPL, p.w - xl));
    q.rowStrideA = p.rowStrideA;
    q.rowStrideB = p.rowStrideB;
    q.vb0 = &vbuf[warp][0][0];
    q.vb1 = &vbuf[warp][1][0];
    q.ring = &pxring[warp][0][0][0];
    if (lane < 4) {  // zero the pad slots once (they feed masked outputs only, but must stay finite)
#pragma unroll
        for (int i = 0; i < CPL; i++) {
#pragma unroll
            for (int v = 0; v < NVB; v++) vbuf[warp][v][i * 36 + 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; j++) q.g[j] = p.g[j];
#pragma unroll
    for (int i = 0; i < CPL; i++) q.valid[i] = (CPL * lane + i < OUTC) && (xl + i < p.w - 8);

    const uint8_t *ia = p.a + (long long)img * p.imgStrideA;
    const uint8_t *ib = p.b + (long long)img * p.imgStrideB;
    {   // Centring constant: luma of image a at the strip's centre pixel (same for every lane).
        int cx = min(X0 + OUTC / 2, p.w - 1), cy = min(Y0 + q.nIn / 2, p.h - 1);
        float f = luma_magic(ld_nc_u32(ia + (long long)cy * p.rowStrideA + (long long)cx * 4));  // 2^23 + L
        q.K = -(f * kLumaScale);                          // a' = F*s + K  ~  luma - luma_c
        q.c = -fmaf(8388608.0f, kLumaScale, q.K);         // the centring this K really applies
    }

    // Warp-uniform choice of the load path.  Lanes entirely right of the image (their outputs are all
    // masked) re-read the strip's first columns so that the whole warp can use aligned vector loads.
    // CPL = 3 copies pixel by pixel (4 bytes, each address clamped into the row): always "fast".
    const bool fast = (CPL == 3 && MODE == 2) || (p.vecOK && __all_sync(0xffffffffu, q.nvalid == CPL || q.nvalid == 0));
    const int xld = (fast && q.nvalid == 0) ? X0 : xl;
    q.pa = ia + (long long)Y0 * p.rowStrideA + (long long)xld * 4;
    q.pb = ib + (long long)Y0 * p.rowStrideB + (long long)xld * 4;
    if (fast && q.nvalid == 0) q.nvalid = CPL;
#pragma unroll
    for (int i = 0; i < CPL; i++) q.pxoff[i] = 4 * min(i, max(q.nvalid, 1) - 1);

    double dsum;
    if (PIPE && CPL == 4) dsum = fast ? walk_strip_pipe(reinterpret_cast<const StripCtx<4> &>(q)) : walk_strip<CPL, false>(q);
    else if (MODE == 2 && CPL == 4) dsum = fast ? walk_strip2<4>(reinterpret_cast<const StripCtx<4> &>(q)) : walk_strip<CPL, false>(q);
    else if (MODE == 2 && CPL == 3) dsum = walk_strip2<3>(reinterpret_cast<const StripCtx<3> &>(q));
    else if (MODE == 3 && CPL == 4) dsum = fast ? walk_stripN<4>(reinterpret_cast<const StripCtx<4> &>(q)) : walk_strip<CPL, false>(q);
    else if (MODE == 4 && CPL == 4) dsum = fast ? walk_stripN<2>(reinterpret_cast<const StripCtx<4> &>(q)) : walk_strip<CPL, false>(q);
    else dsum = fast ? walk_strip<CPL, true>(q) : walk_strip<CPL, false>(q);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    if (lane == 0) p.partials[seg] = dsum * 4.0;
}

// (An explicit register cap was measured with __maxnreg__: 184 registers (11 warps per SM) and 168 (12) both spill and
// lose 7-15 % against the uncapped one-warp blocks; profiles/r2_k1_variants.txt.)

// ------------------------------------------------------------------------------------------------
// Warp-specialised variant (aligned inputs only).  The monolithic kernel above is register-limited to
// 8 warps/SM (its 8-row ring alone is 128 registers) and leaves the FMA pipe ~35 % idle.  Here a CTA is
// two warpgroups: warps 0-3 ("V") keep the ring and do planes + the vertical pass, warps 4-7 ("H") do the
// horizontal pass + SSIM formula + reduction.  V-warp i and H-warp i+4 share a strip and live on the
// same SM sub-partition; rows travel through a 4-deep shared-memory ring guarded by full/empty
// mbarriers.  setmaxnreg moves registers from the H warpgroup (72) to the V warpgroup (184), so two
// such CTAs (16 warps) fit an SM.
// ------------------------------------------------------------------------------------------------
constexpr int kWsPairs = 4;
constexpr int kVStages = 4;
constexpr int kVLanesWs = 36;
constexpr int kVRowF4 = 4 * kVLanesWs;  // float4 per V row

struct __align__(16) WsSmem {
    float4 vrows[kWsPairs][kVStages][kVRowF4];
    uint8_t pxring[kWsPairs][kStages][2][512];
    unsigned long long full[kWsPairs][kVStages];
    unsigned long long empty[kWsPairs][kVStages];
};

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(256, 2) ssim_ws_kernel(const SsimParams p) {
    constexpr int CPL = 4, OUTC = 120;
    extern __shared__ __align__(16) uint8_t ws_raw[];
    WsSmem &sm = *reinterpret_cast<WsSmem *>(ws_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int pair = warp & 3;
    const bool isV = warp < 4;

    if (threadIdx.x < kWsPairs * kVStages) {
        mbar_init(smem_u32(&sm.full[0][0]) + 8u * threadIdx.x, 32);
        mbar_init(smem_u32(&sm.empty[0][0]) + 8u * threadIdx.x, 32);
    }
    // pad slots of the V rows (read by lanes 30/31 for masked outputs) must stay finite
    for (int i = threadIdx.x; i < kWsPairs * kVStages * 4 * 4; i += 256) {
        int row = i >> 4, item = (i >> 2) & 3, padl = i & 3;
        (&sm.vrows[0][0][0])[row * kVRowF4 + item * kVLanesWs + 32 + padl] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const long long segsPerImg = (long long)p.nsx * p.nsy;
    const long long seg = (long long)blockIdx.x * kWsPairs + pair;
    const bool active = seg < segsPerImg * p.n;  // inactive pairs still execute their warpgroup's setmaxnreg
    const int img = active ? (int)(seg / segsPerImg) : 0;
    const int rseg = (int)(seg - (long long)img * segsPerImg);
    const int sy = rseg / p.nsx, sx = rseg - sy * p.nsx;
    const int X0 = sx * OUTC, Y0 = sy * p.rs;
    const int nOut = min(p.rs, (p.h - 8) - Y0);
    const int nIn = nOut + 7;
    const int xl = X0 + CPL * lane;
    const uint8_t *ia = p.a + (long long)img * p.imgStrideA;
    const uint8_t *ib = p.b + (long long)img * p.imgStrideB;
    float K = 0.f, c = 0.f;
    if (active) {   // centring constant, identical in both warps of the pair
        int cx = min(X0 + OUTC / 2, p.w - 1), cy = min(Y0 + nIn / 2, p.h - 1);
        float f = luma_magic(ld_nc_u32(ia + (long long)cy * p.rowStrideA + (long long)cx * 4));
        K = -(f * kLumaScale);
        c = -fmaf(8388608.0f, kLumaScale, K);
    }
    float2 g2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) g2[j] = make_float2(p.g[j], p.g[j]);
    const uint32_t fullBase = smem_u32(&sm.full[pair][0]);
    const uint32_t emptyBase = smem_u32(&sm.empty[pair][0]);
    float4 *vrows = &sm.vrows[pair][0][0];

    if (isV) {
        // ---------------- producer: pixels → planes → vertical 8-tap → V row ring ----------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
        if (!active) return;
        const int xld = (xl + CPL <= p.w) ? xl : X0;  // lanes right of the image re-read the strip start (masked)
        const uint8_t *pa = ia + (long long)Y0 * p.rowStrideA + (long long)xld * 4;
        const uint8_t *pb = ib + (long long)Y0 * p.rowStrideB + (long long)xld * 4;
        const float2 s2 = make_float2(kLumaScale, kLumaScale);
        const float2 K2 = make_float2(K, K);
        const uint32_t myRing = smem_u32(&sm.pxring[pair][0][0][0]) + lane * 16;
        const uint8_t *myRingP = &sm.pxring[pair][0][0][0] + lane * 16;
#pragma unroll
        for (int row = 0; row < kStages; row++) {
            cp_async<16>(myRing + (2 * row) * 512, pa);
            cp_async<16>(myRing + (2 * row + 1) * 512, pb);
            cp_async_commit();
            pa += p.rowStrideA;
            pb += p.rowStrideB;
        }
        float2 rab[8][CPL], rqp[8][CPL];
#define WS_PLANES(S, R)                                                                         \
    {                                                                                           \
        cp_async_wait<kStages - 1>();                                                           \
        const uint8_t *rb_ = myRingP + (2 * ((S) & (kStages - 1))) * 512;                       \
        const uint4 va_ = *reinterpret_cast<const uint4 *>(rb_);                                \
        const uint4 vb_ = *reinterpret_cast<const uint4 *>(rb_ + 512);                          \
        const uint32_t xa_[4] = {va_.x, va_.y, va_.z, va_.w}, xb_[4] = {vb_.x, vb_.y, vb_.z, vb_.w}; \
        _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                       \
            float2 f = make_float2(luma_magic(xa_[i]), luma_magic(xb_[i]));                     \
            float2 t = __ffma2_rn(f, s2, K2);                                                   \
            float2 sq = __fmul2_rn(t, t);                                                       \
            rab[S][i] = t;                                                                      \
            rqp[S][i] = make_float2(sq.x + sq.y, t.x * t.y);                                    \
        }                                                                                       \
        if ((R) + kStages < nIn) {                                                              \
            cp_async<16>(myRing + (2 * ((S) & (kStages - 1))) * 512, pa);                       \
            cp_async<16>(myRing + (2 * ((S) & (kStages - 1)) + 1) * 512, pb);                   \
            pa += p.rowStrideA;                                                                 \
            pb += p.rowStrideB;                                                                 \
        }                                                                                       \
        cp_async_commit();                                                                      \
    }
#pragma unroll
        for (int r = 0; r < 7; r++) WS_PLANES(r, r)

#pragma unroll 1
        for (int r = 7; r < nIn; r++) {
            const int o = r - 7;
            const int st = o & (kVStages - 1);
            mbar_wait(emptyBase + 8u * st, ((o / kVStages) & 1) ^ 1);  // slot free (first lap passes at once)
            float4 *dst = vrows + st * kVRowF4 + lane;
#define WS_STEP(S)                                                                              \
    case S: {                                                                                   \
        WS_PLANES(S, r)                                                                         \
        _Pragma("unroll") for (int i = 0; i < CPL; i++) {                                       \
            float2 vab = __fmul2_rn(rab[(S + 1) & 7][i], g2[0]);                                \
            float2 vqp = __fmul2_rn(rqp[(S + 1) & 7][i], g2[0]);                                \
            _Pragma("unroll") for (int j = 1; j < 8; j++) {                                     \
                vab = __ffma2_rn(rab[(S + 1 + j) & 7][i], g2[j], vab);                          \
                vqp = __ffma2_rn(rqp[(S + 1 + j) & 7][i], g2[j], vqp);                          \
            }                                                                                   \
            dst[i * kVLanesWs] = make_float4(vab.x, vab.y, vqp.x, vqp.y);                       \
        }                                                                                       \
    } break;
            switch (r & 7) {
                WS_STEP(0) WS_STEP(1) WS_STEP(2) WS_STEP(3) WS_STEP(4) WS_STEP(5) WS_STEP(6) WS_STEP(7)
            }
#undef WS_STEP
            mbar_arrive(fullBase + 8u * st);  // release: the row's stores are visible to the consumer
        }
#undef WS_PLANES
        return;
    }

    // ---------------- consumer: horizontal 8-tap + SSIM + reduction ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (!active) return;
    bool valid[CPL];
#pragma unroll
    for (int i = 0; i < CPL; i++) valid[i] = (CPL * lane + i < OUTC) && (xl + i < p.w - 8);
    const float kTh = fmaf(c, c, 0.5f * kC1f);
    const float2 qpInit = make_float2(kC2f, 0.5f * kC2f);
    float fs[CPL];
#pragma unroll
    for (int i = 0; i < CPL; i++) fs[i] = 0.f;
#pragma unroll 1
    for (int o = 0; o < nOut; o++) {
        const int st = o & (kVStages - 1);
        mbar_wait(fullBase + 8u * st, (o / kVStages) & 1);
        const float4 *src = vrows + st * kVRowF4 + lane;
        float2 mab[CPL], mqp[CPL];
#pragma unroll
        for (int i = 0; i < CPL; i++) { mab[i] = make_float2(0.f, 0.f); mqp[i] = qpInit; }
        // scatter form: item k (column 4*lane + k) feeds outputs i = k-7..k with tap k - i
#pragma unroll
        for (int k = 0; k < CPL + 7; k++) {
            const float4 v = src[(k % CPL) * kVLanesWs + k / CPL];
#pragma unroll
            for (int i = 0; i < CPL; i++) {
                const int t = k - i;
                if (t >= 0 && t < 8) {
                    mab[i] = __ffma2_rn(make_float2(v.x, v.y), g2[t], mab[i]);
                    mqp[i] = __ffma2_rn(make_float2(v.z, v.w), g2[t], mqp[i]);
                }
            }
        }
        mbar_arrive(emptyBase + 8u * st);  // all of this lane's reads of the slot are done
#pragma unroll
        for (int i = 0; i < CPL; i++) {
            float m = mab[i].x * mab[i].y;
            float nn = fmaf(mab[i].x, mab[i].x, mab[i].y * mab[i].y);
            float th = fmaf(c, mab[i].x + mab[i].y, kTh);
            float A1h = m + th;
            float B1 = fmaf(2.f, th, nn);
            float A2h = mqp[i].y - m;
            float B2 = mqp[i].x - nn;
            float ssim4 = __fdividef(A1h * A2h, B1 * B2);
            fs[i] += valid[i] ? ssim4 : 0.f;
        }
    }
    double dsum = 0.0;
#pragma unroll
    for (int i = 0; i < CPL; i++) dsum += (double)fs[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    if (lane == 0) p.partials[seg] = dsum * 4.0;
}

// Deterministic per-image reduction of the strip partials → mean SSIM (ssim.go:155-165).
__global__ void ssim_finalize_kernel(const double *partials, int segsPerImg, long long count,
                                     double *scores, long long scoreStride) {
    const int img = blockIdx.x;
    const double *pp = partials + (long long)img * segsPerImg;
    double s = 0.0;
    for (int i = threadIdx.x; i < segsPerImg; i += 32) s += pp[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) scores[(long long)img * scoreStride] = (count == 0) ? 1.0 : s / (double)count;
}

// pixelSSIM (ssim.go:169-204): global statistics in FP64; images here are < 8 px on a side, or the
// caller asked for it explicitly.  One block per pair; two passes like the reference.
__global__ void pixel_ssim_kernel(const uint8_t *a, const uint8_t *b, long long imgStrideA,
                                  long long imgStrideB, int rowStrideA, int rowStrideB, int w, int h,
                                  double *scores, long long scoreStride) {
    __shared__ double red[5][32];
    __shared__ double mu[2];
    const int img = blockIdx.x;
    const uint8_t *ia = a + (long long)img * imgStrideA;
    const uint8_t *ib = b + (long long)img * imgStrideB;
    const long long npx = (long long)w * h;
    const double n = (double)npx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    auto lum = [](const uint8_t *q) {
        return __dadd_rn(__dadd_rn(__dmul_rn(0.299, (double)q[0]), __dmul_rn(0.587, (double)q[1])),
                         __dmul_rn(0.114, (double)q[2]));
    };
    auto block_sum = [&](double v, int slot) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[slot][warp] = v;
    };
    double sa = 0, sb = 0;
    for (long long i = threadIdx.x; i < npx; i += blockDim.x) {
        int y = (int)(i / w), x = (int)(i - (long long)y * w);
        sa += lum(ia + (long long)y * rowStrideA + x * 4);
        sb += lum(ib + (long long)y * rowStrideB + x * 4);
    }
    block_sum(sa, 0);
    block_sum(sb, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0, tb = 0;
        for (int k = 0; k < nw; k++) { ta += red[0][k]; tb += red[1][k]; }
        mu[0] = ta / n;
        mu[1] = tb / n;
    }
    __syncthreads();
    const double muA = mu[0], muB = mu[1];
    double saa = 0, sbb = 0, sab = 0;
    for (long long i = threadIdx.x; i < npx; i += blockDim.x) {
        int y = (int)(i / w), x = (int)(i - (long long)y * w);
        double da = lum(ia + (long long)y * rowStrideA + x * 4) - muA;
        double db = lum(ib + (long long)y * rowStrideB + x * 4) - muB;
        saa += da * da;
        sbb += db * db;
        sab += da * db;
    }
    block_sum(saa, 2);
    block_sum(sbb, 3);
    block_sum(sab, 4);
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 1.0;
        if (npx > 0) {
            double taa = 0, tbb = 0, tab = 0;
            for (int k = 0; k < nw; k++) { taa += red[2][k]; tbb += red[3][k]; tab += red[4][k]; }
            taa /= n; tbb /= n; tab /= n;
            double num = (2 * muA * muB + 6.5025) * (2 * tab + 58.5225);
            double den = (muA * muA + muB * muB + 6.5025) * (taa + tbb + 58.5225);
            r = num / den;
        }
        scores[(long long)img * scoreStride] = r;
    }
}

// MSSSIM tail (ssim.go:349-364): out = exp(sum_i w_i * ln(max(score_i, 1e-10))).
__global__ void msssim_combine_kernel(const double *levelScores, int nLevels, int n,
                                      const double *weights, double *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = 0.0;
    for (int l = 0; l < nLevels; l++) r += weights[l] * log(fmax(levelScores[(long long)l * n + i], 1e-10));   // level-major
    out[i] = exp(r);
}

void gaussian1d(float g[8]) {
    // 1-D factor of ssim.go:223-241: exp(-(x^2+y^2)/(2*1.5^2)) / sum == g(x)*g(y)
    double v[8], s = 0.0;
    for (int k = -4; k < 4; k++) { v[k + 4] = exp(-(double)(k * k) / 4.5); s += v[k + 4]; }
    for (int k = 0; k < 8; k++) g[k] = (float)(v[k] / s);
}

#ifndef FB_SSIM_CPL_DEFAULT
#define FB_SSIM_CPL_DEFAULT 4
#endif
// Segment geometry: strips of OUTC outputs; rs rows per segment chosen so the grid fills the GPU.
struct Geo { int cpl, outc, nsx, nsy, rs; };
int ssim_cpl() {
    static int cpl = [] {
        const char *e = getenv("FB_SSIM_CPL");  // tuning knob: columns per lane (2, 3 or 4)
        return (e && e[0] == '2') ? 2 : (e && e[0] == '3') ? 3 : (e && e[0] == '4') ? 4 : FB_SSIM_CPL_DEFAULT;
    }();
    return cpl;
}
Geo geometry(int w, int h, int n) {
    Geo g;
    g.cpl = ssim_cpl();
    g.outc = 32 * g.cpl - 8;
    int ow = w - 8, oh = h - 8;
    g.nsx = (ow + g.outc - 1) / g.outc;
    // Rows per segment: every segment pays a 7-row warm-up (work ~ oh + 7*oh/rs per strip) and the last blocks of the
    // grid straggle for about half a segment.  Measured on 16 4K pairs: rs 64 / 128 / 256 / 512 -> 0.546 / 0.542 /
    // 0.574 / 0.612 ms; a closed-form optimum of that model (rs = 112 at 16 pairs, 160 at 32) was no better than a
    // flat 128 (0.551 / 1.053 vs 0.542 / 1.050 ms), so: 128, halved (down to 16) while the grid has fewer than 12
    // warps per SM (one 4K pair: 32 rows, 0.050 ms against 0.056 ms at 128; 16 MS-SSIM thumbnails: 16 rows).
    // FB_SSIM_RS overrides for experiments.
    static const int rsForce = [] { const char *e = getenv("FB_SSIM_RS"); int v = e ? atoi(e) : 0; return (v >= 16 && v <= 4096) ? v : 0; }();
    const long long want = 148LL * 12;
    int rs = rsForce ? rsForce : 128;
    if (!rsForce)
        while (rs > 16 && (long long)n * g.nsx * ((oh + rs - 1) / rs) < want) rs >>= 1;
    g.rs = rs;
    g.nsy = (oh + rs - 1) / rs;
    return g;
}

}  // namespace

size_t ssim_scratch_bytes(int w, int h, int n) {
    if (w < 9 || h < 9) return 256;
    Geo g = geometry(w, h, n);
    return align_up(sizeof(double) * (size_t)n * g.nsx * g.nsy, 256);
}

// The strip kernels read pixels with cp.async.cg / ld.global.nc.L1::no_allocate: L1 has nothing to cache, so the whole
// L1/shared array goes to shared memory (one-warp blocks need 10 x 13 KB; the driver's default carve-out heuristic is
// free to pick less).  FB_SSIM_DEBUG=1 prints the resident blocks per SM once.
template <typename K>
static void prepare_kernel(K kernel, int threads, const char *name) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    static const bool dbg = getenv("FB_SSIM_DEBUG") != nullptr;
    if (dbg) {
        int nb = 0;
        cudaFuncAttributes fa;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, 0);
        cudaFuncGetAttributes(&fa, kernel);
        fprintf(stderr, "[fb] %s: %d regs, %zu B static smem, %d blocks (%d warps) per SM\n", name, fa.numRegs, fa.sharedSizeBytes, nb, nb * threads / 32);
    }
}
#define FB_PREPARE(kernel, threads)                                          \
    do {                                                                     \
        static bool done_ = false;                                           \
        if (!done_) { prepare_kernel(kernel, threads, #kernel); done_ = true; } \
    } while (0)

// {w, h, n} tensor map over a batch of NRGBA images (pixels as 32-bit elements), boxes of 128 pixels x kTmaRows rows.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}
static bool make_image_map(CUtensorMap *m, const uint8_t *base, long long imgStride, int rowStride, int w, int h, int n, int boxRows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t img = imgStride > 0 ? (cuuint64_t)imgStride : (cuuint64_t)rowStride * (cuuint64_t)h;
    const cuuint64_t strides[2] = {(cuuint64_t)rowStride, img};
    const cuuint32_t box[3] = {128u, (cuuint32_t)boxRows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Scores for n equal-sized pairs.  Dispatch of ssim.go:35-42: w<8||h<8 → pixelSSIM, else windowed.
int launch_ssim(DevCtx *c, cudaStream_t s, const uint8_t *a, const uint8_t *b, long long imgStrideA,
                long long imgStrideB, int rowStrideA, int rowStrideB, int w, int h, int n,
                double *scores, long long scoreStride, void *scratch) {
    (void)c;
    if (n <= 0) return FB_OK;
    if (w < 8 || h < 8) {
        pixel_ssim_kernel<<<n, 256, 0, s>>>(a, b, imgStrideA, imgStrideB, rowStrideA, rowStrideB, w, h,
                                            scores, scoreStride);
        FB_LAUNCHED(1);
        FB_CUDA(cudaGetLastError());
        return FB_OK;
    }
    if (w == 8 || h == 8) {  // zero windows → 1.0 (ssim.go:162-164)
        ssim_finalize_kernel<<<n, 32, 0, s>>>((const double *)scratch, 0, 0, scores, scoreStride);
        FB_LAUNCHED(1);
        FB_CUDA(cudaGetLastError());
        return FB_OK;
    }
    Geo g = geometry(w, h, n);
    SsimParams p;
    p.a = a; p.b = b;
    p.imgStrideA = imgStrideA; p.imgStrideB = imgStrideB;
    p.rowStrideA = rowStrideA; p.rowStrideB = rowStrideB;
    p.w = w; p.h = h; p.n = n;
    p.nsx = g.nsx; p.nsy = g.nsy; p.rs = g.rs;
    p.vecOK = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)imgStrideA | (uintptr_t)imgStrideB |
                (uintptr_t)rowStrideA | (uintptr_t)rowStrideB) & 15) == 0;
    p.partials = (double *)scratch;
    gaussian1d(p.g);
    long long segs = (long long)n * g.nsx * g.nsy;
    long long blocks = (segs + 3) / 4;
    static const bool useWs = [] { const char *e = getenv("FB_SSIM_WS"); return e && e[0] == '1'; }();
    if (useWs && g.cpl == 4 && p.vecOK && (w & 3) == 0) {
        static bool attrSet = false;
        if (!attrSet) {
            FB_CUDA(cudaFuncSetAttribute(ssim_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WsSmem)));
            attrSet = true;
        }
        ssim_ws_kernel<<<(unsigned)blocks, 256, sizeof(WsSmem), s>>>(p);
    } else if (g.cpl == 4) {
        // FB_SSIM_PIPE=1 / FB_SSIM_MODE=0|1|2 pick the row walk, FB_SSIM_WPB=1|4 the warps per block (experiments;
        // every combination is parity-tested in tests/test_variants_gpu.py).
        static const int mode = [] {
            const char *e = getenv("FB_SSIM_PIPE");
            if (e && e[0] == '1') return 1;
            const char *m = getenv("FB_SSIM_MODE");
            return (m && m[0] >= '0' && m[0] <= '6') ? m[0] - '0' : FB_SSIM_MODE_DEFAULT;
        }();
        static const int wpb = [] { const char *e = getenv("FB_SSIM_WPB"); return (e && e[0] == '1') ? 1 : (e && e[0] == '4') ? 4 : FB_SSIM_WPB_DEFAULT; }();
#define FB_LAUNCH_STRIP(MODE_, WPB_)                                                           \
    do {                                                                                          \
        FB_PREPARE((ssim_strip_kernel<4, MODE_, WPB_>), 32 * WPB_);                                \
        ssim_strip_kernel<4, MODE_, WPB_><<<(unsigned)((segs + WPB_ - 1) / WPB_), 32 * WPB_, 0, s>>>(p); \
    } while (0)
        bool done = false;
        if ((mode == 5 || mode == 6) && p.vecOK && (imgStrideA > 0 || n == 1) && (imgStrideB > 0 || n == 1)) {   // TMA-staged rows
            CUtensorMap ta, tb;
            const int boxRows = mode == 5 ? 4 : 8;
            if (make_image_map(&ta, a, imgStrideA, rowStrideA, w, h, n, boxRows) && make_image_map(&tb, b, imgStrideB, rowStrideB, w, h, n, boxRows)) {
                if (mode == 5) {
                    FB_PREPARE((ssim_strip_tma_kernel<4, 4>), 32);
                    ssim_strip_tma_kernel<4, 4><<<(unsigned)segs, 32, 0, s>>>(ta, tb, p);
                } else {
                    FB_PREPARE((ssim_strip_tma_kernel<8, 2>), 32);
                    ssim_strip_tma_kernel<8, 2><<<(unsigned)segs, 32, 0, s>>>(ta, tb, p);
                }
                done = true;
            }
        }
        if (done) {}
        else if (mode >= 5) FB_LAUNCH_STRIP(2, 1);   // unaligned buffers / no driver entry point: the cp.async twin
        else if (mode == 1) FB_LAUNCH_STRIP(1, 4);
        else if (mode == 3) FB_LAUNCH_STRIP(3, 1);
        else if (mode == 4) FB_LAUNCH_STRIP(4, 1);
        else if (mode == 2 && wpb == 1) FB_LAUNCH_STRIP(2, 1);
        else if (mode == 2) FB_LAUNCH_STRIP(2, 2);   // 4 V buffers: 2 warps stay under 48 KB static
        else if (wpb == 1) FB_LAUNCH_STRIP(0, 1);
        else FB_LAUNCH_STRIP(0, 4);
#undef FB_LAUNCH_STRIP
    }
    else if (g.cpl == 3) {
        FB_PREPARE((ssim_strip_kernel<3, 2, 1>), 32);
        ssim_strip_kernel<3, 2, 1><<<(unsigned)segs, 32, 0, s>>>(p);
    }
    else ssim_strip_kernel<2><<<(unsigned)blocks, 128, 0, s>>>(p);
    FB_CUDA(cudaGetLastError());
    ssim_finalize_kernel<<<n, 32, 0, s>>>(p.partials, g.nsx * g.nsy, (long long)(w - 8) * (h - 8), scores,
                                          scoreStride);
    FB_CUDA(cudaGetLastError());
    FB_LAUNCHED(2);
    return FB_OK;
}

int launch_pixel_ssim(cudaStream_t s, const uint8_t *a, const uint8_t *b, long long imgStrideA,
                      long long imgStrideB, int rowStrideA, int rowStrideB, int w, int h, int n,
                      double *scores, long long scoreStride) {
    if (n <= 0) return FB_OK;
    pixel_ssim_kernel<<<n, 256, 0, s>>>(a, b, imgStrideA, imgStrideB, rowStrideA, rowStrideB, w, h, scores,
                                        scoreStride);
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

int launch_msssim_combine(cudaStream_t s, const double *levelScores, int nLevels, int n,
                          const double *weights_dev, double *out) {
    msssim_combine_kernel<<<(n + 127) / 128, 128, 0, s>>>(levelScores, nLevels, n, weights_dev, out);
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

}  // namespace fb
