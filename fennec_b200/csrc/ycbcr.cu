// ycbcr.cu — SURVEY §8(f1): convertToNRGBA (convert.go:34-64) for what jpeg.Decode returns, on the device.
//
// The quality search decodes every candidate JPEG (compress.go:53-59) into an *image.YCbCr and the reference
// then walks img.At(x,y).RGBA() per pixel.  Uploading the planes (1.5 B/px at 4:2:0 instead of 4 B/px NRGBA) and
// converting here removes that loop and 63 % of the PCIe bytes of every search iteration.
//
// Arithmetic = Go's color.YCbCr.RGBA() followed by convert.go:48-53's `uint8(v >> 8)` (image/color/ycbcr.go, Go
// 1.25.5 standard library — not under /root/reference; restated):
//     yy1 = Y * 0x10101;  r = yy1 + 91881*cr1;  g = yy1 - 22554*cb1 - 46802*cr1;  b = yy1 + 116130*cb1
//     channel = (v in [0, 2^24)) ? v >> 16 : (v < 0 ? 0 : 255)            ==  clamp(v >> 16, 0, 255)
// (arithmetic shift; the equality is checked exhaustively over all 2^24 triples in tests/).  Chroma addressing is
// (*image.YCbCr).COffset with Rect.Min == (0,0).  HBM-bound: 1.5 B read + 4 B written per pixel at 4:2:0.
//
// Mapping: one thread = 4 adjacent pixels of one row: one 32-bit Y load, 1-4 chroma bytes per plane, one 128-bit
// store; the second row of a 4:2:0 pair re-reads its chroma from L1/L2.  Grid = (ceil(w/4/128), h, n).
#include "common.cuh"

namespace fb {

namespace {

struct YccParams {
    const uint8_t *y, *cb, *cr;
    uint8_t *dst;
    long long yImgStride, cImgStride, dstImgStride;
    int yStride, cStride, dstRowStride;
    int w, h;
    int xShift, yShift;  // chroma subsampling as shifts: x >> xShift, y >> yShift
    int vecOK;           // Y rows 4-byte aligned and dst rows 16-byte aligned
};

__device__ __forceinline__ uint32_t ycc_px(int Y, int cb, int cr) {
    const int yy1 = Y * 0x10101, cb1 = cb - 128, cr1 = cr - 128;
    const int r = min(max((yy1 + 91881 * cr1) >> 16, 0), 255);
    const int g = min(max((yy1 - 22554 * cb1 - 46802 * cr1) >> 16, 0), 255);
    const int b = min(max((yy1 + 116130 * cb1) >> 16, 0), 255);
    return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | 0xFF000000u;
}

__global__ void __launch_bounds__(128) ycbcr_to_nrgba_kernel(const YccParams p) {
    const int x0 = (blockIdx.x * 128 + threadIdx.x) * 4;
    const int y = blockIdx.y, img = blockIdx.z;
    if (x0 >= p.w) return;
    const uint8_t *yrow = p.y + (long long)img * p.yImgStride + (long long)y * p.yStride;
    const long long coff = (long long)img * p.cImgStride + (long long)(y >> p.yShift) * p.cStride;
    const uint8_t *cbrow = p.cb + coff, *crrow = p.cr + coff;
    uint8_t *drow = p.dst + (long long)img * p.dstImgStride + (long long)y * p.dstRowStride + (long long)x0 * 4;
    if (p.vecOK && x0 + 4 <= p.w) {
        const uint32_t y4 = __ldg(reinterpret_cast<const uint32_t *>(yrow + x0));
        uint32_t out[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int cx = (x0 + i) >> p.xShift;
            out[i] = ycc_px((int)((y4 >> (8 * i)) & 0xFF), (int)__ldg(cbrow + cx), (int)__ldg(crrow + cx));
        }
        *reinterpret_cast<uint4 *>(drow) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {
        for (int i = 0; i < 4 && x0 + i < p.w; i++) {
            const int cx = (x0 + i) >> p.xShift;
            *reinterpret_cast<uint32_t *>(drow + 4 * i) = ycc_px((int)__ldg(yrow + x0 + i), (int)__ldg(cbrow + cx), (int)__ldg(crrow + cx));
        }
    }
}

// 4:2:0 fast path: one thread = 8 pixels x 2 rows = one 4-sample chroma group: two 64-bit Y loads, one 32-bit
// load per chroma plane, four 128-bit stores (the generic kernel above issues 9 loads per 4 pixels).
__global__ void __launch_bounds__(128) ycbcr420_to_nrgba_kernel(const YccParams p) {
    const int x0 = (blockIdx.x * 128 + threadIdx.x) * 8;
    const int y0 = blockIdx.y * 2, img = blockIdx.z;
    if (x0 >= p.w) return;
    const uint8_t *yrow = p.y + (long long)img * p.yImgStride + (long long)y0 * p.yStride;
    const long long coff = (long long)img * p.cImgStride + (long long)blockIdx.y * p.cStride;
    const uint8_t *cbrow = p.cb + coff, *crrow = p.cr + coff;
    uint8_t *drow = p.dst + (long long)img * p.dstImgStride + (long long)y0 * p.dstRowStride + (long long)x0 * 4;
    const bool two = y0 + 1 < p.h;
    if (x0 + 8 <= p.w) {
        const uint32_t cb4 = __ldg(reinterpret_cast<const uint32_t *>(cbrow + (x0 >> 1)));
        const uint32_t cr4 = __ldg(reinterpret_cast<const uint32_t *>(crrow + (x0 >> 1)));
        const uint2 ya = __ldg(reinterpret_cast<const uint2 *>(yrow + x0));
        const uint2 yb = two ? __ldg(reinterpret_cast<const uint2 *>(yrow + p.yStride + x0)) : make_uint2(0u, 0u);
        uint32_t oa[8], ob[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int cb = (int)((cb4 >> (8 * (i >> 1))) & 0xFF), cr = (int)((cr4 >> (8 * (i >> 1))) & 0xFF);
            const int cb1 = cb - 128, cr1 = cr - 128;
            const int dr = 91881 * cr1, dg = -22554 * cb1 - 46802 * cr1, db = 116130 * cb1;   // shared by the two rows
            const int Ya = (int)(((i < 4 ? ya.x : ya.y) >> (8 * (i & 3))) & 0xFF) * 0x10101;
            const int Yb = (int)(((i < 4 ? yb.x : yb.y) >> (8 * (i & 3))) & 0xFF) * 0x10101;
            oa[i] = (uint32_t)min(max((Ya + dr) >> 16, 0), 255) | ((uint32_t)min(max((Ya + dg) >> 16, 0), 255) << 8) |
                    ((uint32_t)min(max((Ya + db) >> 16, 0), 255) << 16) | 0xFF000000u;
            ob[i] = (uint32_t)min(max((Yb + dr) >> 16, 0), 255) | ((uint32_t)min(max((Yb + dg) >> 16, 0), 255) << 8) |
                    ((uint32_t)min(max((Yb + db) >> 16, 0), 255) << 16) | 0xFF000000u;
        }
        *reinterpret_cast<uint4 *>(drow) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
        *reinterpret_cast<uint4 *>(drow + 16) = make_uint4(oa[4], oa[5], oa[6], oa[7]);
        if (two) {
            *reinterpret_cast<uint4 *>(drow + p.dstRowStride) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
            *reinterpret_cast<uint4 *>(drow + p.dstRowStride + 16) = make_uint4(ob[4], ob[5], ob[6], ob[7]);
        }
    } else {  // right edge: per pixel
        for (int r = 0; r < (two ? 2 : 1); r++)
            for (int i = 0; x0 + i < p.w; i++) {
                const int cx = (x0 + i) >> 1;
                *reinterpret_cast<uint32_t *>(drow + (long long)r * p.dstRowStride + 4 * i) =
                    ycc_px((int)__ldg(yrow + (long long)r * p.yStride + x0 + i), (int)__ldg(cbrow + cx), (int)__ldg(crrow + cx));
            }
    }
}

__global__ void __launch_bounds__(128) gray_to_nrgba_kernel(const uint8_t *g, long long gImgStride, int gStride, uint8_t *dst,
                                                            long long dstImgStride, int dstRowStride, int w, int h, int vecOK) {
    const int x0 = (blockIdx.x * 128 + threadIdx.x) * 4;
    const int y = blockIdx.y, img = blockIdx.z;
    if (x0 >= w) return;
    const uint8_t *grow = g + (long long)img * gImgStride + (long long)y * gStride;
    uint8_t *drow = dst + (long long)img * dstImgStride + (long long)y * dstRowStride + (long long)x0 * 4;
    if (vecOK && x0 + 4 <= w) {
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(grow + x0));
        uint32_t out[4];
#pragma unroll
        for (int i = 0; i < 4; i++) out[i] = ((v >> (8 * i)) & 0xFF) * 0x010101u | 0xFF000000u;
        *reinterpret_cast<uint4 *>(drow) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {
        for (int i = 0; i < 4 && x0 + i < w; i++)
            *reinterpret_cast<uint32_t *>(drow + 4 * i) = (uint32_t)__ldg(grow + x0 + i) * 0x010101u | 0xFF000000u;
    }
}

}  // namespace

// ratio: Go's image.YCbCrSubsampleRatio constant (444, 422, 420, 440, 411, 410).
bool ycbcr_ratio_shifts(int ratio, int *xShift, int *yShift) {
    static const int xs[6] = {0, 1, 1, 0, 2, 2}, ys[6] = {0, 0, 1, 1, 0, 1};
    if (ratio < 0 || ratio > 5) return false;
    *xShift = xs[ratio];
    *yShift = ys[ratio];
    return true;
}

int launch_ycbcr_to_nrgba(cudaStream_t s, const uint8_t *y, long long yImgStride, int yStride, const uint8_t *cb,
                          const uint8_t *cr, long long cImgStride, int cStride, int w, int h, int ratio, uint8_t *dst,
                          long long dstImgStride, int dstRowStride, int n) {
    if (n <= 0 || w <= 0 || h <= 0) return FB_OK;
    YccParams p;
    if (!ycbcr_ratio_shifts(ratio, &p.xShift, &p.yShift)) return FB_E_INVALID;
    p.y = y; p.cb = cb; p.cr = cr; p.dst = dst;
    p.yImgStride = yImgStride; p.cImgStride = cImgStride; p.dstImgStride = dstImgStride;
    p.yStride = yStride; p.cStride = cStride; p.dstRowStride = dstRowStride;
    p.w = w; p.h = h;
    p.vecOK = ((((uintptr_t)y | (uintptr_t)yImgStride | (uintptr_t)yStride) & 3) == 0) &&
              ((((uintptr_t)dst | (uintptr_t)dstImgStride | (uintptr_t)dstRowStride) & 15) == 0);
    const bool fast420 = ratio == 2 && ((((uintptr_t)y | (uintptr_t)yImgStride | (uintptr_t)yStride) & 7) == 0) &&
                         ((((uintptr_t)cb | (uintptr_t)cr | (uintptr_t)cImgStride | (uintptr_t)cStride) & 3) == 0) &&
                         ((((uintptr_t)dst | (uintptr_t)dstImgStride | (uintptr_t)dstRowStride) & 15) == 0);
    if (fast420) {
        dim3 grid(((w + 7) / 8 + 127) / 128, (h + 1) / 2, n);
        ycbcr420_to_nrgba_kernel<<<grid, 128, 0, s>>>(p);
    } else {
        dim3 grid(((w + 3) / 4 + 127) / 128, h, n);
        ycbcr_to_nrgba_kernel<<<grid, 128, 0, s>>>(p);
    }
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

int launch_gray_to_nrgba(cudaStream_t s, const uint8_t *g, long long gImgStride, int gStride, int w, int h, uint8_t *dst,
                         long long dstImgStride, int dstRowStride, int n) {
    if (n <= 0 || w <= 0 || h <= 0) return FB_OK;
    const int vecOK = ((((uintptr_t)g | (uintptr_t)gImgStride | (uintptr_t)gStride) & 3) == 0) &&
                      ((((uintptr_t)dst | (uintptr_t)dstImgStride | (uintptr_t)dstRowStride) & 15) == 0);
    dim3 grid(((w + 3) / 4 + 127) / 128, h, n);
    gray_to_nrgba_kernel<<<grid, 128, 0, s>>>(g, gImgStride, gStride, dst, dstImgStride, dstRowStride, w, h, vecOK);
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

}  // namespace fb
