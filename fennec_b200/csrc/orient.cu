// orient.cu — SURVEY §8(f4): ApplyOrientation (exif.go:176-203) = the rotate / flip loops of convert.go:186-256
// on a device-resident NRGBA image.  Pure pixel permutations: 4 B read + 4 B written per pixel, HBM-bound.
//
// With (c, r) a destination column / row and w, h the SOURCE dims, following the reference loops literally:
//   2 FlipH       D(c, r) = S(w-1-c, r)                          convert.go:229-241
//   3 Rotate180   D(c, r) = S(w-1-c, h-1-r)                      convert.go:201-213
//   4 FlipV       D(c, r) = S(c, h-1-r)                          convert.go:244-255
//   5 Transpose   rotate270CW then FlipH (exif.go:187-190)  →   D(c, r) = S(w-1-r, h-1-c)     (dst is h x w)
//   6 Rotate90CW  D(c, r) = S(r, h-1-c)                          convert.go:186-198
//   7 Transverse  rotate90CW then FlipH (exif.go:193-196)   →   D(c, r) = S(r, c)
//   8 Rotate270CW D(c, r) = S(w-1-r, c)                          convert.go:216-226
// Orientations 2-4 keep rows as rows: one thread per pixel, both sides coalesced (a reversed warp still covers one
// 128-byte segment).  Orientations 5-8 exchange the axes: 32x32 tiles through shared memory (33-word pitch), read
// along source rows, written along destination rows.
#include "common.cuh"

namespace fb {

namespace {

struct OrientParams {
    const uint8_t *src;
    uint8_t *dst;
    long long srcImgStride, dstImgStride;
    int srcRowStride, dstRowStride;
    int w, h;          // source dims
    int flipX, flipY;  // applied to SOURCE coordinates
};

__global__ void __launch_bounds__(256) orient_rows_kernel(const OrientParams p) {
    const int c = blockIdx.x * 256 + threadIdx.x, r = blockIdx.y, img = blockIdx.z;
    if (c >= p.w) return;
    const int sx = p.flipX ? p.w - 1 - c : c, sy = p.flipY ? p.h - 1 - r : r;
    const uint32_t v = ld_nc_u32(p.src + (long long)img * p.srcImgStride + (long long)sy * p.srcRowStride + (long long)sx * 4);
    *reinterpret_cast<uint32_t *>(p.dst + (long long)img * p.dstImgStride + (long long)r * p.dstRowStride + (long long)c * 4) = v;
}

// Same mapping, 4 pixels per thread with 128-bit accesses (w % 4 == 0, 16-byte aligned rows).
__global__ void __launch_bounds__(256) orient_rows_vec_kernel(const OrientParams p) {
    const int c = (blockIdx.x * 256 + threadIdx.x) * 4, r = blockIdx.y, img = blockIdx.z;
    if (c >= p.w) return;
    const int sy = p.flipY ? p.h - 1 - r : r;
    const int sx = p.flipX ? p.w - 4 - c : c;
    uint4 v = ld_nc_u128(p.src + (long long)img * p.srcImgStride + (long long)sy * p.srcRowStride + (long long)sx * 4);
    if (p.flipX) v = make_uint4(v.w, v.z, v.y, v.x);
    *reinterpret_cast<uint4 *>(p.dst + (long long)img * p.dstImgStride + (long long)r * p.dstRowStride + (long long)c * 4) = v;
}

// D(c, r) = S(fx(r), fy(c)): destination is h wide, w tall.
__global__ void __launch_bounds__(256) orient_transpose_kernel(const OrientParams p) {
    __shared__ uint32_t tile[32][33];
    const int img = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8 threads
    const int sx0 = blockIdx.x * 32, sy0 = blockIdx.y * 32;          // source tile origin
    const uint8_t *s = p.src + (long long)img * p.srcImgStride;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int sx = sx0 + tx, sy = sy0 + ty + 8 * k;
        if (sx < p.w && sy < p.h) tile[ty + 8 * k][tx] = ld_nc_u32(s + (long long)sy * p.srcRowStride + (long long)sx * 4);
    }
    __syncthreads();
    uint8_t *d = p.dst + (long long)img * p.dstImgStride;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int sy = sy0 + tx, sx = sx0 + ty + 8 * k;              // lanes run along the source column = destination row
        if (sx < p.w && sy < p.h) {
            const int c = p.flipY ? p.h - 1 - sy : sy, r = p.flipX ? p.w - 1 - sx : sx;
            *reinterpret_cast<uint32_t *>(d + (long long)r * p.dstRowStride + (long long)c * 4) = tile[tx][ty + 8 * k];
        }
    }
}

}  // namespace

// Destination dims of ApplyOrientation; returns false for orientations that return the input (1, 0, unknown).
bool orient_dims(int orient, int w, int h, int *dw, int *dh) {
    if (orient < 2 || orient > 8) return false;
    const bool swap = orient >= 5;
    *dw = swap ? h : w;
    *dh = swap ? w : h;
    return true;
}

int launch_orient(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int w, int h, int orient,
                  uint8_t *dst, long long dstImgStride, int dstRowStride, int n) {
    if (n <= 0 || w <= 0 || h <= 0) return FB_OK;
    OrientParams p;
    p.src = src; p.dst = dst;
    p.srcImgStride = srcImgStride; p.dstImgStride = dstImgStride;
    p.srcRowStride = srcRowStride; p.dstRowStride = dstRowStride;
    p.w = w; p.h = h;
    switch (orient) {
        case 2: p.flipX = 1; p.flipY = 0; break;
        case 3: p.flipX = 1; p.flipY = 1; break;
        case 4: p.flipX = 0; p.flipY = 1; break;
        case 5: p.flipX = 1; p.flipY = 1; break;   // D(c, r) = S(w-1-r, h-1-c)
        case 6: p.flipX = 0; p.flipY = 1; break;   // D(c, r) = S(r, h-1-c)
        case 7: p.flipX = 0; p.flipY = 0; break;   // D(c, r) = S(r, c)
        case 8: p.flipX = 1; p.flipY = 0; break;   // D(c, r) = S(w-1-r, c)
        default: return FB_E_INVALID;
    }
    const bool vec = (w & 3) == 0 && ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)srcImgStride | (uintptr_t)dstImgStride |
                                        (uintptr_t)srcRowStride | (uintptr_t)dstRowStride) & 15) == 0);
    if (orient <= 4 && vec) orient_rows_vec_kernel<<<dim3((w / 4 + 255) / 256, h, n), 256, 0, s>>>(p);
    else if (orient <= 4) orient_rows_kernel<<<dim3((w + 255) / 256, h, n), 256, 0, s>>>(p);
    else orient_transpose_kernel<<<dim3((w + 31) / 32, (h + 31) / 32, n), 256, 0, s>>>(p);
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

}  // namespace fb
