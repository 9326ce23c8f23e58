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

template <bool VERTICAL>
int launch_pass(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, uint8_t *dst,
                long long dstImgStride, int dstRowStride, int outW, int outH, int n, const int *start,
                const int *index, const double *weight) {
    if (n <= 0 || outW <= 0 || outH <= 0) return FB_OK;
    ResizeParams p;
    p.src = src; p.dst = dst;
    p.srcImgStride = srcImgStride; p.dstImgStride = dstImgStride;
    p.srcRowStride = srcRowStride; p.dstRowStride = dstRowStride;
    p.outW = outW; p.outH = outH;
    p.start = start; p.index = index; p.weight = weight;
    dim3 grid((outW + 255) / 256, outH, n);
    resize_pass_kernel<VERTICAL><<<grid, 256, 0, s>>>(p);
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

}  // namespace

int launch_resize_h(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int srcW,
                    int srcH, uint8_t *dst, long long dstImgStride, int dstRowStride, int dstW, int n,
                    const int *start_dev, const int *index_dev, const double *weight_dev, int maxTaps) {
    (void)srcW; (void)maxTaps;
    return launch_pass<false>(s, src, srcImgStride, srcRowStride, dst, dstImgStride, dstRowStride, dstW, srcH, n,
                              start_dev, index_dev, weight_dev);
}

int launch_resize_v(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int srcW,
                    int srcH, uint8_t *dst, long long dstImgStride, int dstRowStride, int dstH, int n,
                    const int *start_dev, const int *index_dev, const double *weight_dev, int maxTaps) {
    (void)srcH; (void)maxTaps;
    return launch_pass<true>(s, src, srcImgStride, srcRowStride, dst, dstImgStride, dstRowStride, srcW, dstH, n,
                             start_dev, index_dev, weight_dev);
}

}  // namespace fb
