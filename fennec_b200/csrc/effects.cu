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

}  // namespace

int launch_gaussian_blur(cudaStream_t s, const uint8_t *src, uint8_t *dst, long long imgStride,
                         int rowStride, int w, int h, int n, const double *kernel_dev,
                         const float *kernel32_dev, int radius, uint8_t *tmp, long long tmpImgStride,
                         int tmpRowStride) {
    if (n <= 0 || w <= 0 || h <= 0) return FB_OK;
    BlurParams p;
    p.w = w; p.h = h; p.radius = radius;
    p.kernel = kernel_dev; p.kernel32 = kernel32_dev;
    p.eps = (float)((2 * radius + 3) * 255.0 * 1.1920928955078125e-07);
    p.exactOnly = (p.eps >= 0.25f) ? 1 : 0;  // absurdly long kernels: no useful fast path
    dim3 grid((w + 255) / 256, h, n);
    // horizontal: src → tmp
    p.src = src; p.srcImgStride = imgStride; p.srcRowStride = rowStride;
    p.alpha = src; p.alphaImgStride = imgStride; p.alphaRowStride = rowStride;
    p.dst = tmp; p.dstImgStride = tmpImgStride; p.dstRowStride = tmpRowStride;
    blur_pass_kernel<false><<<grid, 256, 0, s>>>(p);
    // vertical: tmp → dst, alpha from the original
    p.src = tmp; p.srcImgStride = tmpImgStride; p.srcRowStride = tmpRowStride;
    p.dst = dst; p.dstImgStride = imgStride; p.dstRowStride = rowStride;
    blur_pass_kernel<true><<<grid, 256, 0, s>>>(p);
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
    dim3 grid((w + 255) / 256, h, n);
    sharpen_kernel<<<grid, 256, 0, s>>>(p);
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
