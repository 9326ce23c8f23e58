// pixfmt.cu — convertToNRGBA (convert.go:34-64) for the concrete types image/png and image/jpeg decode into besides
// YCbCr and Gray (ycbcr.cu): *image.RGBA, *image.RGBA64, *image.NRGBA64, *image.Gray16, *image.CMYK, *image.Paletted.
// The reference walks img.At(x, y).RGBA() through the color.Color interface per pixel; here the decoded Pix buffer
// is uploaded as it is (1-8 B/px) and converted on the device, so the NRGBA image the rest of the path works on
// never exists on the host.
//
// Arithmetic, per pixel, all uint32 as in Go (image/color/color.go, image/color/ycbcr.go — Go 1.25.5 standard
// library, not under /root/reference; restated):
//   step 1  At().RGBA():  RGBA     c | c << 8 per field            RGBA64  the big-endian 16-bit fields
//                         NRGBA64  c * a / 0xffff, a               Gray16  (y, y, y, 0xffff)
//                         CMYK     (0xffff - c*0x101) * (0xffff - k*0x101) / 0xffff, alpha 0xffff
//                         Paletted Palette[index].RGBA() (passed in as 4 x uint16 per entry)
//   step 2  convert.go:42-60:  a == 0 -> 0;  a == 0xffff -> c >> 8, 255;  else uint8(((c * 0xffff) / a) >> 8), a >> 8
//           — uint8() keeps the low byte, which matters for c > a (not a valid premultiplied colour).
// Integer only, bit-exact.  HBM-bound: bpp bytes read + 4 written per pixel; the un-premultiply division is taken only
// by warps that hold a translucent pixel.
//
// Mapping: one thread = 4 adjacent pixels (one or two 128-bit loads, or a 64- / 32-bit one; one 128-bit store), 8 rows
// per block so the Paletted table (256 entries converted once per block into shared memory) is amortised.
#include "common.cuh"

namespace fb {

namespace {

constexpr int kPfThreads = 256;
constexpr int kPfRows = 8;

struct PixFmtParams {
    const uint8_t *src;
    uint8_t *dst;
    const uint16_t *pal16;      // Paletted: per image 256 entries x 4 uint16 (r, g, b, a as Color.RGBA() returns them)
    unsigned int *badIndex;     // Paletted: set to 1 when an index >= ncolors is met (Go panics there); may be null
    long long srcImgStride, dstImgStride;
    int srcRowStride, dstRowStride;
    int w, h, ncolors;
    int vecOK;                  // src rows aligned for the format's vector load, dst rows 16-byte aligned
};

// convert.go:42-60 on one pixel's 16-bit (r, g, b, a) → packed R | G<<8 | B<<16 | A<<24
__device__ __forceinline__ uint32_t convert_px(uint32_t r, uint32_t g, uint32_t b, uint32_t a) {
    if (a == 0u) return 0u;
    if (a == 0xFFFFu) return ((r >> 8) & 0xFFu) | (((g >> 8) & 0xFFu) << 8) | (((b >> 8) & 0xFFu) << 16) | 0xFF000000u;
    const uint32_t R = ((r * 0xFFFFu) / a) >> 8, G = ((g * 0xFFFFu) / a) >> 8, B = ((b * 0xFFFFu) / a) >> 8;
    return (R & 0xFFu) | ((G & 0xFFu) << 8) | ((B & 0xFFu) << 16) | ((a >> 8) << 24);
}

// big-endian 16-bit field k (0..3) of an 8-byte pixel held as two little-endian words
__device__ __forceinline__ uint32_t be16(uint32_t lo, uint32_t hi, int k) {
    const uint32_t wd = (k < 2) ? lo : hi;
    const uint32_t half = (k & 1) ? (wd >> 16) : (wd & 0xFFFFu);     // bytes [msb, lsb] in memory order = lsb-first in the word
    return ((half & 0xFFu) << 8) | (half >> 8);
}

template <int FMT>
__device__ __forceinline__ uint32_t px_from_words(uint32_t lo, uint32_t hi) {
    if (FMT == 1) {            // RGBA 8-bit premultiplied
        const uint32_t r = lo & 0xFFu, g = (lo >> 8) & 0xFFu, b = (lo >> 16) & 0xFFu, a = lo >> 24;
        if (a == 0xFFu) return lo;                                   // (c | c << 8) >> 8 == c
        return convert_px(r * 0x101u, g * 0x101u, b * 0x101u, a * 0x101u);
    } else if (FMT == 2) {     // RGBA64
        return convert_px(be16(lo, hi, 0), be16(lo, hi, 1), be16(lo, hi, 2), be16(lo, hi, 3));
    } else if (FMT == 3) {     // NRGBA64
        const uint32_t a = be16(lo, hi, 3);
        return convert_px(be16(lo, hi, 0) * a / 0xFFFFu, be16(lo, hi, 1) * a / 0xFFFFu, be16(lo, hi, 2) * a / 0xFFFFu, a);
    } else if (FMT == 4) {     // Gray16: lo = the two bytes [msb, lsb]
        const uint32_t y = lo & 0xFFu;                               // (y16 >> 8) = the first byte in memory
        return y * 0x010101u | 0xFF000000u;
    } else {                   // CMYK
        const uint32_t wk = 0xFFFFu - (lo >> 24) * 0x101u;
        const uint32_t r = (0xFFFFu - (lo & 0xFFu) * 0x101u) * wk / 0xFFFFu;
        const uint32_t g = (0xFFFFu - ((lo >> 8) & 0xFFu) * 0x101u) * wk / 0xFFFFu;
        const uint32_t b = (0xFFFFu - ((lo >> 16) & 0xFFu) * 0x101u) * wk / 0xFFFFu;
        return (r >> 8) | ((g >> 8) << 8) | ((b >> 8) << 16) | 0xFF000000u;
    }
}

template <int FMT> struct FmtBpp { static constexpr int v = (FMT == 2 || FMT == 3) ? 8 : (FMT == 4 ? 2 : (FMT == 6 ? 1 : 4)); };

__device__ __forceinline__ uint32_t ld_bytes_le(const uint8_t *p, int n) {      // n <= 4 bytes, any alignment
    uint32_t v = 0;
    for (int i = 0; i < n; i++) v |= (uint32_t)__ldg(p + i) << (8 * i);
    return v;
}

template <int FMT>
__global__ void __launch_bounds__(kPfThreads) pixfmt_to_nrgba_kernel(const PixFmtParams p) {
    constexpr int BPP = FmtBpp<FMT>::v;
    __shared__ uint32_t table[256];
    const int img = blockIdx.z;
    if (FMT == 6) {
        const uint16_t *pal = p.pal16 + (size_t)img * 1024;
        for (int i = threadIdx.x; i < 256; i += kPfThreads) {
            uint32_t v = 0u;
            if (i < p.ncolors) {
                const uint2 e = __ldg(reinterpret_cast<const uint2 *>(pal) + i);        // r | g << 16, b | a << 16
                v = convert_px(e.x & 0xFFFFu, e.x >> 16, e.y & 0xFFFFu, e.y >> 16);
            }
            table[i] = v;
        }
        __syncthreads();
    }
    const int x0 = (blockIdx.x * kPfThreads + threadIdx.x) * 4;
    if (x0 >= p.w) return;
    const int yEnd = min((int)(blockIdx.y + 1) * kPfRows, p.h);
    bool bad = false;
    for (int y = blockIdx.y * kPfRows; y < yEnd; y++) {
        const uint8_t *s = p.src + (long long)img * p.srcImgStride + (long long)y * p.srcRowStride + (long long)x0 * BPP;
        uint8_t *d = p.dst + (long long)img * p.dstImgStride + (long long)y * p.dstRowStride + (long long)x0 * 4;
        uint32_t o[4];
        const bool full = x0 + 4 <= p.w;
        if (full && p.vecOK) {
            if (BPP == 4) {
                const uint4 q = ld_nc_u128(s);
                o[0] = px_from_words<FMT>(q.x, 0u); o[1] = px_from_words<FMT>(q.y, 0u);
                o[2] = px_from_words<FMT>(q.z, 0u); o[3] = px_from_words<FMT>(q.w, 0u);
            } else if (BPP == 8) {
                const uint4 q0 = ld_nc_u128(s), q1 = ld_nc_u128(s + 16);
                o[0] = px_from_words<FMT>(q0.x, q0.y); o[1] = px_from_words<FMT>(q0.z, q0.w);
                o[2] = px_from_words<FMT>(q1.x, q1.y); o[3] = px_from_words<FMT>(q1.z, q1.w);
            } else if (BPP == 2) {
                const uint2 q = ld_nc_u64(s);
                o[0] = px_from_words<FMT>(q.x & 0xFFFFu, 0u); o[1] = px_from_words<FMT>(q.x >> 16, 0u);
                o[2] = px_from_words<FMT>(q.y & 0xFFFFu, 0u); o[3] = px_from_words<FMT>(q.y >> 16, 0u);
            } else {
                const uint32_t q = ld_nc_u32(s);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t i = (q >> (8 * k)) & 0xFFu;
                    bad |= (int)i >= p.ncolors;
                    o[k] = table[i];
                }
            }
            *reinterpret_cast<uint4 *>(d) = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
            for (int k = 0; k < 4 && x0 + k < p.w; k++) {
                uint32_t v;
                if (BPP == 8) v = px_from_words<FMT>(ld_bytes_le(s + 8 * k, 4), ld_bytes_le(s + 8 * k + 4, 4));
                else if (BPP == 1) {
                    const uint32_t i = __ldg(s + k);
                    bad |= (int)i >= p.ncolors;
                    v = table[i];
                } else v = px_from_words<FMT>(ld_bytes_le(s + BPP * k, BPP), 0u);
                uint8_t *dk = d + 4 * k;
                if ((((uintptr_t)dk) & 3) == 0) *reinterpret_cast<uint32_t *>(dk) = v;
                else { dk[0] = (uint8_t)v; dk[1] = (uint8_t)(v >> 8); dk[2] = (uint8_t)(v >> 16); dk[3] = (uint8_t)(v >> 24); }
            }
        }
    }
    if (FMT == 6 && bad && p.badIndex) *p.badIndex = 1u;      // benign race: every writer stores 1
}

}  // namespace

int pixfmt_bytes_per_pixel(int fmt) {
    switch (fmt) {
        case 1: case 5: return 4;
        case 2: case 3: return 8;
        case 4: return 2;
        case 6: return 1;
        default: return 0;
    }
}

int launch_pixfmt_to_nrgba(cudaStream_t s, int fmt, const uint8_t *src, long long srcImgStride, int srcRowStride, int w, int h,
                           const uint16_t *pal16_dev, int ncolors, uint8_t *dst, long long dstImgStride, int dstRowStride, int n,
                           unsigned int *badIndex_dev) {
    const int bpp = pixfmt_bytes_per_pixel(fmt);
    if (!bpp) return FB_E_INVALID;
    if (n <= 0 || w <= 0 || h <= 0) return FB_OK;
    if (fmt == 6 && (!pal16_dev || ncolors < 1 || ncolors > 256)) return FB_E_INVALID;
    PixFmtParams p;
    p.src = src; p.dst = dst; p.pal16 = pal16_dev; p.badIndex = badIndex_dev;
    p.srcImgStride = srcImgStride; p.dstImgStride = dstImgStride;
    p.srcRowStride = srcRowStride; p.dstRowStride = dstRowStride;
    p.w = w; p.h = h; p.ncolors = ncolors;
    const uintptr_t srcAlign = (uintptr_t)(4 * bpp > 16 ? 16 : 4 * bpp) - 1;     // bytes of 4 pixels, at most one 128-bit load each
    p.vecOK = ((((uintptr_t)src | (uintptr_t)srcImgStride | (uintptr_t)srcRowStride) & srcAlign) == 0) &&
              ((((uintptr_t)dst | (uintptr_t)dstImgStride | (uintptr_t)dstRowStride) & 15) == 0);
    const dim3 grid(((w + 3) / 4 + kPfThreads - 1) / kPfThreads, (h + kPfRows - 1) / kPfRows, n);
    switch (fmt) {
        case 1: pixfmt_to_nrgba_kernel<1><<<grid, kPfThreads, 0, s>>>(p); break;
        case 2: pixfmt_to_nrgba_kernel<2><<<grid, kPfThreads, 0, s>>>(p); break;
        case 3: pixfmt_to_nrgba_kernel<3><<<grid, kPfThreads, 0, s>>>(p); break;
        case 4: pixfmt_to_nrgba_kernel<4><<<grid, kPfThreads, 0, s>>>(p); break;
        case 5: pixfmt_to_nrgba_kernel<5><<<grid, kPfThreads, 0, s>>>(p); break;
        default: pixfmt_to_nrgba_kernel<6><<<grid, kPfThreads, 0, s>>>(p); break;
    }
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

}  // namespace fb
