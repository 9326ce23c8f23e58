// palette.cu — SURVEY §8(f3): applyPalette + palettedToNRGBA (targetsize.go:479-545), the per-pixel half of
// target-size strategy 2 (median-cut quantisation; medianCut itself samples <= 100 000 pixels and sorts on the host).
//
// applyPalette: index of the palette colour with the smallest squared RGB distance, FIRST minimum on ties
// (`dist < bestDist` scanning i upwards, targetsize.go:499-510; the map there is only a memo).  Entries are NRGBA with
// A = 255 (medianCut's average(), targetsize.go:400-413), so c.RGBA()>>8 is the 8-bit channel.
//
// Per (pixel, entry): dot = dp4a(x, p) (alpha lane zeroed), key = (|p|^2 * 256 + i) - 512 * dot — i.e.
// (|p|^2 - 2 x.p) * 256 + i, which orders entries by (distance, index) lexicographically because |x|^2 is common
// to all entries of a pixel — and a signed min: 3 integer instructions.  Not HBM-bound: 256 entries cost ~770
// integer ops per pixel against 9 bytes of traffic; the roofline it is measured against is the integer pipe.
// One thread = 4 pixels (128-bit load, 32-bit index store, optional 128-bit NRGBA store); the palette lives in
// shared memory and every lane reads the same entry (broadcast).
#include "common.cuh"

namespace fb {

namespace {

struct PalParams {
    const uint8_t *src;
    uint8_t *idx;       // w x h indices (image.Paletted.Pix), may be null
    uint8_t *out;       // NRGBA reconstruction (palettedToNRGBA), may be null
    const uint8_t *palettes;   // per image: 256 entries x 4 bytes (R, G, B, A)
    long long srcImgStride, idxImgStride, outImgStride;
    int srcRowStride, idxRowStride, outRowStride;
    int w, h, ncolors;
    int vecOK;
};

__global__ void __launch_bounds__(256) apply_palette_kernel(const PalParams p) {
    __shared__ uint32_t pal[256];     // R | G<<8 | B<<16
    __shared__ int base[256];         // |p|^2 * 256 + i
    __shared__ uint32_t palOut[256];  // R | G<<8 | B<<16 | A<<24 as palettedToNRGBA writes it
    const int img = blockIdx.z;
    for (int i = threadIdx.x; i < p.ncolors; i += 256) {
        const uint32_t e = __ldg(reinterpret_cast<const uint32_t *>(p.palettes + (size_t)img * 1024) + i);
        const int r = e & 0xFF, g = (e >> 8) & 0xFF, b = (e >> 16) & 0xFF;
        pal[i] = e & 0x00FFFFFFu;
        base[i] = (r * r + g * g + b * b) * 256 + i;
        palOut[i] = e;
    }
    __syncthreads();
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= p.w) return;
    const uint8_t *row = p.src + (long long)img * p.srcImgStride + (long long)y * p.srcRowStride + (long long)x0 * 4;
    uint32_t px[4];
    const bool full = p.vecOK && x0 + 4 <= p.w;
    if (full) {
        const uint4 q = ld_nc_u128(row);
        px[0] = q.x; px[1] = q.y; px[2] = q.z; px[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) px[k] = (x0 + k < p.w) ? ld_nc_u32(row + 4 * k) : 0u;
    }
    int best[4] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF};
    uint32_t rgb[4];
#pragma unroll
    for (int k = 0; k < 4; k++) rgb[k] = px[k] & 0x00FFFFFFu;
#pragma unroll 4
    for (int i = 0; i < p.ncolors; i++) {
        const uint32_t e = pal[i];
        const int bi = base[i];
#pragma unroll
        for (int k = 0; k < 4; k++) best[k] = min(best[k], bi - 512 * (int)__dp4a(rgb[k], e, 0u));
    }
    uint32_t packed = 0;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t bi = (uint32_t)best[k] & 0xFFu;
        packed |= bi << (8 * k);
        o[k] = palOut[bi];
    }
    if (p.idx) {
        uint8_t *ip = p.idx + (long long)img * p.idxImgStride + (long long)y * p.idxRowStride + x0;
        if (x0 + 4 <= p.w && (((uintptr_t)ip) & 3) == 0) *reinterpret_cast<uint32_t *>(ip) = packed;
        else
            for (int k = 0; k < 4 && x0 + k < p.w; k++) ip[k] = (uint8_t)(packed >> (8 * k));
    }
    if (p.out) {
        uint8_t *op = p.out + (long long)img * p.outImgStride + (long long)y * p.outRowStride + (long long)x0 * 4;
        if (x0 + 4 <= p.w && (((uintptr_t)op) & 15) == 0) *reinterpret_cast<uint4 *>(op) = make_uint4(o[0], o[1], o[2], o[3]);
        else
            for (int k = 0; k < 4 && x0 + k < p.w; k++) *reinterpret_cast<uint32_t *>(op + 4 * k) = o[k];
    }
}

}  // namespace

int launch_apply_palette(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int w, int h,
                         const uint8_t *palettes_dev, int ncolors, uint8_t *idx, long long idxImgStride, int idxRowStride,
                         uint8_t *out, long long outImgStride, int outRowStride, int n) {
    if (n <= 0 || w <= 0 || h <= 0) return FB_OK;
    if (ncolors < 1 || ncolors > 256) return FB_E_INVALID;
    PalParams p;
    p.src = src; p.idx = idx; p.out = out; p.palettes = palettes_dev;
    p.srcImgStride = srcImgStride; p.idxImgStride = idxImgStride; p.outImgStride = outImgStride;
    p.srcRowStride = srcRowStride; p.idxRowStride = idxRowStride; p.outRowStride = outRowStride;
    p.w = w; p.h = h; p.ncolors = ncolors;
    p.vecOK = (((uintptr_t)src | (uintptr_t)srcImgStride | (uintptr_t)srcRowStride) & 15) == 0;
    apply_palette_kernel<<<dim3(((w + 3) / 4 + 255) / 256, h, n), 256, 0, s>>>(p);
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

}  // namespace fb
