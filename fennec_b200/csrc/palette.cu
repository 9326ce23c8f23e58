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
//
// Cell lists (images of >= kCellMinPixels): the brute-force scan is 85 % wasted — almost every entry is provably
// farther than some other one for ALL colours near the pixel.  RGB space is cut into 32^3 cells of 8^3 colours; per
// palette a pre-kernel (one warp per cell) computes, for every entry p, dmin(p) / dmax(p) = squared distance from p
// to the nearest / farthest point of the cell, U = min_p dmax(p), and keeps the entries with dmin(p) <= U in
// ascending index order.  For any colour x of the cell the reference's winner i* (first minimum) satisfies
// dmin(i*) <= |x - p_i*|^2 <= |x - p_q|^2 <= dmax(q) for every q, so i* is on the list, and the (distance, index)
// key picks it among the listed entries exactly as among all of them — ties included.  A list that would exceed 31
// entries (a palette crowded into a few cells) marks the cell kCellOverflow and its pixels take the full scan, so
// the result never depends on the lists, only the time does.  Typical median-cut palettes list 4-12 entries per
// cell: ~60 integer instructions per pixel instead of ~770, and the kernel moves from the integer pipe to HBM/L2.
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
    const uint4 *cellList;     // per image: kCells records of 32 bytes: [count | kCellOverflow][<= 31 entry indices, ascending]
};

#ifndef FB_PAL_ROWS
#define FB_PAL_ROWS 8
#endif
constexpr int kCellShift = 3;                       // 8 colours per axis per cell
constexpr int kCellAxis = 256 >> kCellShift;        // 32
constexpr int kCells = kCellAxis * kCellAxis * kCellAxis;
constexpr int kCellCap = 31;                        // + the count byte = one 32-byte record, two 16-byte loads at most
constexpr int kCellOverflow = 255;
constexpr long long kCellMinPixels = 1 << 17;       // below this the 8.4 M (cell, entry) tests cost more than they save

// Candidates of a cell = entries whose nearest point of the cell is no farther than the best worst case.  One warp
// per run of 8 cells along the R axis: the G and B terms of dmin / dmax are shared by the run and computed once per
// entry (8 entries per lane), each cell then adds its R term, reduces U over the warp and compacts the survivors in
// index order with ballots (a round with no survivor costs three instructions).
__global__ void __launch_bounds__(256) palette_cells_kernel(const uint8_t *palettes, int ncolors, uint8_t *cellList) {
    __shared__ uint8_t rec[8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int run = blockIdx.x * 8 + warp, img = blockIdx.y;            // kCells / 8 runs
    const uint32_t *pal = reinterpret_cast<const uint32_t *>(palettes + (size_t)img * 1024);
    constexpr int runsPerRow = kCellAxis / 8;
    const int cx0 = (run % runsPerRow) * 8, cy = (run / runsPerRow) % kCellAxis, cz = run / (runsPerRow * kCellAxis);
    const int lo1 = cy << kCellShift, lo2 = cz << kCellShift;
    constexpr int span = (1 << kCellShift) - 1;
    constexpr int kFar = 0x3FFFFFFF;                                     // entries past ncolors: never near, never best
    int c0[8], nn[8], ff[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int i = r * 32 + lane;
        c0[r] = 0; nn[r] = kFar; ff[r] = kFar;
        if (i < ncolors) {
            const uint32_t e = __ldg(pal + i);
            const int c1 = (e >> 8) & 0xFF, c2 = (e >> 16) & 0xFF;
            const int a1 = c1 - lo1, a2 = c2 - lo2;
            const int n1 = max(max(-a1, a1 - span), 0), f1 = max(a1, span - a1);
            const int n2 = max(max(-a2, a2 - span), 0), f2 = max(a2, span - a2);
            c0[r] = e & 0xFF;
            nn[r] = n1 * n1 + n2 * n2;
            ff[r] = f1 * f1 + f2 * f2;
        }
    }
    for (int j = 0; j < 8; j++) {
        const int lo0 = (cx0 + j) << kCellShift;
        int dmin[8], umin = 0x7FFFFFFF;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int a0 = c0[r] - lo0;
            const int n0 = max(max(-a0, a0 - span), 0), f0 = max(a0, span - a0);
            dmin[r] = nn[r] + n0 * n0;
            umin = min(umin, ff[r] + f0 * f0);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) umin = min(umin, __shfl_xor_sync(0xFFFFFFFFu, umin, o));
        int total = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const bool cand = dmin[r] <= umin;      // umin < kFar because ncolors >= 1
            const unsigned m = __ballot_sync(0xFFFFFFFFu, cand);
            if (m) {
                const int pos = total + __popc(m & ((1u << lane) - 1u));
                if (cand && pos < kCellCap) rec[warp][1 + pos] = (uint8_t)(r * 32 + lane);
                total += __popc(m);
            }
        }
        if (lane == 0) rec[warp][0] = (uint8_t)(total > kCellCap ? kCellOverflow : total);
        __syncwarp();
        const uint8_t v = lane <= total ? rec[warp][lane] : rec[warp][1];      // tail: repeats the first entry (see the matcher)
        const int cell = (cx0 + j) | (cy << 5) | (cz << 10);
        cellList[((size_t)img * kCells + cell) * 32 + lane] = v;
        __syncwarp();
    }
}

// ROWS image rows per block: the three 1 KB palette tables are staged once per block, and with the cell lists a row
// costs less than that staging.
template <bool CELLS, int ROWS>
__global__ void __launch_bounds__(256) apply_palette_kernel(const PalParams p) {
    __shared__ uint2 palBase[256];    // x: R | G<<8 | B<<16;  y: |p|^2 * 256 + i — one 64-bit LDS per (pixel, entry)
    __shared__ uint32_t palOut[256];  // R | G<<8 | B<<16 | A<<24 as palettedToNRGBA writes it
    const int img = blockIdx.z;
    for (int i = threadIdx.x; i < p.ncolors; i += 256) {
        const uint32_t e = __ldg(reinterpret_cast<const uint32_t *>(p.palettes + (size_t)img * 1024) + i);
        const int r = e & 0xFF, g = (e >> 8) & 0xFF, b = (e >> 16) & 0xFF;
        palBase[i] = make_uint2(e & 0x00FFFFFFu, (uint32_t)((r * r + g * g + b * b) * 256 + i));
        palOut[i] = e;
    }
    __syncthreads();
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x0 >= p.w) return;
    const bool full = p.vecOK && x0 + 4 <= p.w;
    const int yBeg = blockIdx.y * ROWS, yEnd = min(yBeg + ROWS, p.h);
    const uint8_t *row = p.src + (long long)img * p.srcImgStride + (long long)yBeg * p.srcRowStride + (long long)x0 * 4;
    auto load_px = [&](const uint8_t *r, uint32_t (&px)[4]) {
        if (full) {
            const uint4 q = ld_nc_u128(r);
            px[0] = q.x; px[1] = q.y; px[2] = q.z; px[3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) px[k] = (x0 + k < p.w) ? ld_nc_u32(r + 4 * k) : 0u;
        }
    };
    uint32_t nxt[4];
    load_px(row, nxt);
    for (int y = yBeg; y < yEnd; y++) {
        uint32_t rgb[4];
#pragma unroll
        for (int k = 0; k < 4; k++) rgb[k] = nxt[k] & 0x00FFFFFFu;
        row += p.srcRowStride;
        if (y + 1 < yEnd) load_px(row, nxt);      // the next row's pixels travel while this row is matched
        int best[4] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF};
        if (CELLS) {
            const uint4 *lists = p.cellList + (size_t)img * kCells * 2;
            int cell[4];
            uint4 r0[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t c = rgb[k];
                cell[k] = ((c & 0xFF) >> kCellShift) | ((((c >> 8) & 0xFF) >> kCellShift) << 5) | (((c >> 16) >> kCellShift) << 10);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) r0[k] = __ldg(lists + cell[k] * 2);     // count + 15 entries; all four in flight
            bool overflow = false;
            int cnt[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                cnt[k] = r0[k].x & 0xFF;
                if (cnt[k] == kCellOverflow) { overflow = true; continue; }
                const uint32_t c = rgb[k];
                const int n = cnt[k];
                int b = 0x7FFFFFFF;
                // list position q (0-based) sits in byte q + 1 of the record; bytes past the count repeat the first
                // entry, so whole words are matched without a per-entry test
                auto entry = [&](uint32_t wd, int byte) {
                    const uint2 pb = palBase[(wd >> (8 * byte)) & 0xFF];
                    b = min(b, (int)pb.y - 512 * (int)__dp4a(c, pb.x, 0u));
                };
                auto word = [&](uint32_t wd) { entry(wd, 0); entry(wd, 1); entry(wd, 2); entry(wd, 3); };
                entry(r0[k].x, 1); entry(r0[k].x, 2); entry(r0[k].x, 3);
                if (n > 3) word(r0[k].y);
                if (n > 7) word(r0[k].z);
                if (n > 11) word(r0[k].w);
                if (n > 15) {
                    const uint4 r1 = __ldg(lists + cell[k] * 2 + 1);
                    word(r1.x);
                    if (n > 19) word(r1.y);
                    if (n > 23) word(r1.z);
                    if (n > 27) word(r1.w);
                }
                best[k] = b;
            }
            if (overflow) {                       // crowded cell: the full scan, for the pixels that need it
                for (int i = 0; i < p.ncolors; i++) {
                    const uint2 pb = palBase[i];
                    const uint32_t e = pb.x;
                    const int bi = (int)pb.y;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (cnt[k] == kCellOverflow) best[k] = min(best[k], bi - 512 * (int)__dp4a(rgb[k], e, 0u));
                }
            }
        } else {
#pragma unroll 4
            for (int i = 0; i < p.ncolors; i++) {
                const uint2 pb = palBase[i];
                const uint32_t e = pb.x;
                const int bi = (int)pb.y;
#pragma unroll
                for (int k = 0; k < 4; k++) best[k] = min(best[k], bi - 512 * (int)__dp4a(rgb[k], e, 0u));
            }
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
}

}  // namespace

bool palette_uses_cells(int w, int h) {
    static const int mode = [] { const char *e = getenv("FB_PALETTE_CELLS"); return e ? atoi(e) : -1; }();   // 0 / 1 force
    if (mode >= 0) return mode != 0;
    return (long long)w * h >= kCellMinPixels;
}

size_t palette_scratch_bytes(int w, int h, int n) {
    return palette_uses_cells(w, h) ? (size_t)n * kCells * 32 + 256 : 0;
}

int launch_apply_palette(cudaStream_t s, const uint8_t *src, long long srcImgStride, int srcRowStride, int w, int h,
                         const uint8_t *palettes_dev, int ncolors, uint8_t *idx, long long idxImgStride, int idxRowStride,
                         uint8_t *out, long long outImgStride, int outRowStride, int n, void *scratch) {
    if (n <= 0 || w <= 0 || h <= 0) return FB_OK;
    if (ncolors < 1 || ncolors > 256) return FB_E_INVALID;
    PalParams p;
    p.src = src; p.idx = idx; p.out = out; p.palettes = palettes_dev;
    p.srcImgStride = srcImgStride; p.idxImgStride = idxImgStride; p.outImgStride = outImgStride;
    p.srcRowStride = srcRowStride; p.idxRowStride = idxRowStride; p.outRowStride = outRowStride;
    p.w = w; p.h = h; p.ncolors = ncolors;
    p.vecOK = (((uintptr_t)src | (uintptr_t)srcImgStride | (uintptr_t)srcRowStride) & 15) == 0;
    constexpr int rowsCells = FB_PAL_ROWS;      // 2..32 measured flat on B200 (the kernel is L1-data-pipe bound)
    const dim3 grid(((w + 3) / 4 + 255) / 256, h, n), gridCells(grid.x, (h + rowsCells - 1) / rowsCells, n);
    p.cellList = nullptr;
    if (scratch && palette_uses_cells(w, h)) {
        palette_cells_kernel<<<dim3(kCells / 64, n), 256, 0, s>>>(palettes_dev, ncolors, (uint8_t *)scratch);
        FB_LAUNCHED(1);
        p.cellList = (const uint4 *)scratch;
        apply_palette_kernel<true, rowsCells><<<gridCells, 256, 0, s>>>(p);
    } else if (h <= 65535) {
        apply_palette_kernel<false, 1><<<grid, 256, 0, s>>>(p);
    } else {                                      // gridDim.y limit: a very tall, narrow image
        apply_palette_kernel<false, rowsCells><<<gridCells, 256, 0, s>>>(p);
    }
    FB_LAUNCHED(1);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
}

}  // namespace fb
