// Pipe-throughput microbenchmarks for sm_100a (B200): what the SSIM / blur / Lanczos kernels are
// bounded by when they are not HBM-bound.  Prints lane-ops per clock per SM for each instruction
// mix.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) bench(float *out, int iters, float seed) {
    float a[16];
    double d[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i] = seed + i; u[i] = threadIdx.x * 2654435761u + i; }
    float w = seed * 0.5f, c = seed * 0.25f;
    double dw = seed * 0.5, dc = seed * 0.25;
    __shared__ float4 sm[256 * 2];
    sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    sm[threadIdx.x + 256] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {  // FFMA x16
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(w), "f"(c));
        } else if (MODE == 1) {  // FFMA2 x8 (16 lane-FMAs)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned long long p, q, r;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(q) : "f"(w));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(q), "l"(r));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p));
            }
        } else if (MODE == 2) {  // DFMA x8
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dw), "d"(dc));
        } else if (MODE == 3) {  // DADD+DMUL x4 each
#pragma unroll
            for (int i = 0; i < 4; i++) {
                asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(dw));
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i + 4]) : "d"(dc));
            }
        } else if (MODE == 4) {  // I2F from byte x8
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned b = (u[i] >> 8) & 0xff;
                float f;
                asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(b));
                a[i] += f;
                u[i] += 0x01010101u;
            }
        } else if (MODE == 5) {  // PRMT x8
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
        } else if (MODE == 6) {  // dp4a x8
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(0x01020304));
        } else if (MODE == 7) {  // IMAD x8
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(12345));
        } else if (MODE == 8) {  // LDS.128 x8
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float4 v = sm[(threadIdx.x + i * 32 + (int)a[15] * 0) & 511];
                a[i] += v.x + v.y + v.z + v.w;
            }
        } else if (MODE == 9) {  // MUFU.RCP x8
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        } else if (MODE == 10) {  // FFMA2 x8 + I2F x4 co-issue
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned long long p, q, r;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(q) : "f"(w));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(q), "l"(r));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p));
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float f;
                asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(u[i] & 0xff));
                c += f;
                u[i] += 0x01010101u;
            }
        } else if (MODE == 11) {  // FFMA2 x8 + DFMA x4 co-issue
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned long long p, q, r;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(q) : "f"(w));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(q), "l"(r));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p));
            }
#pragma unroll
            for (int i = 0; i < 4; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dw), "d"(dc));
        } else if (MODE == 12) {  // FFMA x16 + LDS.128 x2
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(w), "f"(c));
#pragma unroll
            for (int i = 0; i < 2; i++) {
                float4 v = sm[(threadIdx.x + i * 32 + it) & 511];
                c += v.x;
            }
        } else if (MODE == 13) {  // FADD2 x8
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned long long p, q;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(q) : "f"(w));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p));
            }
        } else if (MODE == 14) {  // SHFL x8
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __shfl_down_sync(0xffffffffu, a[i], 1);
        } else if (MODE == 15) {  // I2F.F64 from int x4 + DADD
#pragma unroll
            for (int i = 0; i < 4; i++) { d[i] += (double)(int)(u[i] & 0xff); u[i] += 0x01010101u; }
        } else if (MODE == 17) {  // dp2a lo+hi x4 each (8 IDP.2A)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                unsigned t;
                asm volatile("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(114u), "r"(u[i]), "r"(0x4B000000u));
                asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(u[i + 4]) : "r"(299u | (587u << 16)), "r"(u[i]), "r"(t));
                u[i] += u[i + 4];
            }
        } else if (MODE == 18 || MODE == 19 || MODE == 20) {  // FFMA2 x8 (+ FFMA x8 | + IDP x8)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned long long p, q, r;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(q) : "f"(w));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(q), "l"(r));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p));
                if (MODE == 19) { float t = __uint_as_float(u[i]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(t) : "f"(w), "f"(c)); u[i] = __float_as_uint(t); }
                if (MODE == 20) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(0x01020304));
            }
        } else if (MODE == 16) {  // HFMA2 x8 (fp16x2)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(0x3c003c00), "r"(0x00010001));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; i++) s += (float)d[i] + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + c;
}

// The SSIM kernel's vertical-pass pattern: 8 accumulator pairs, 8 taps, distinct x pairs, scalar weights.
template <int PACKED>
__global__ void __launch_bounds__(128) filt(float *out, int iters, float seed) {
    float2 x[8][4];
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        g[j] = seed * (j + 1) * 0.01f;
#pragma unroll
        for (int i = 0; i < 4; i++) x[j][i] = make_float2(seed + j + threadIdx.x, seed - i);
    }
    float2 tot = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float2 acc = make_float2(0.f, 0.f), acc2 = make_float2(1.f, 1.f);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (PACKED) {
                    acc = __ffma2_rn(x[j][i], make_float2(g[j], g[j]), acc);
                    acc2 = __ffma2_rn(x[(j + 3) & 7][i], make_float2(g[j], g[j]), acc2);
                } else {
                    acc.x = fmaf(x[j][i].x, g[j], acc.x); acc.y = fmaf(x[j][i].y, g[j], acc.y);
                    acc2.x = fmaf(x[(j + 3) & 7][i].x, g[j], acc2.x); acc2.y = fmaf(x[(j + 3) & 7][i].y, g[j], acc2.y);
                }
            }
            tot.x += acc.x + acc2.x; tot.y += acc.y + acc2.y;
            x[i][i].x += tot.x * 1e-9f;  // keep the ring live and changing (static index: stays in registers)
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot.x + tot.y;
}

template <int PACKED>
static int run_filt(const char *name, float *out, int sms, int clk_khz, int blocksPerSM) {
    int blocks = sms * blocksPerSM;
    filt<PACKED><<<blocks, 128>>>(out, 16, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); filt<PACKED><<<blocks, 128>>>(out, ITERS, 1.0f); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double total = 128.0 /*lane fma per iter*/ * ITERS * 128.0 * blocks;
    double per_s = total / (best * 1e-3);
    printf("%-34s %8.3f ms  %7.2f lane-fma/clk/SM (%d warps/SM)\n", name, best, per_s / sms / (clk_khz * 1e3), blocksPerSM * 4);
    return 0;
}

struct Case { const char *name; int mode; double lane_ops_per_iter; };

template <int MODE>
static int run(const char *name, double ops, float *out, int sms, int clk_khz) {
    int blocks = sms * 4;
    bench<MODE><<<blocks, 256>>>(out, 64, 1.0f);
    CHECK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        bench<MODE><<<blocks, 256>>>(out, ITERS, 1.0f);
        cudaEventRecord(e1);
        CHECK(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double total = ops * ITERS * 256.0 * blocks;
    double per_s = total / (best * 1e-3);
    printf("%-28s %8.3f ms  %9.2f Gop/s  %7.2f lane-ops/clk/SM @max-clk(%d MHz)\n", name, best, per_s * 1e-9,
           per_s / sms / (clk_khz * 1e3), clk_khz / 1000);
    return 0;
}

int main() {
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s SMs=%d clk=%d kHz\n", p.name, p.multiProcessorCount, clk);
    float *out;
    CHECK(cudaMalloc(&out, sizeof(float) * 256 * p.multiProcessorCount * 4));
    int sms = p.multiProcessorCount;
    run<0>("FFMA x16", 16, out, sms, clk);
    run<1>("FFMA2 x8 (16 lane-fma)", 16, out, sms, clk);
    run<13>("FADD2 x8 (16 lane-add)", 16, out, sms, clk);
    run<2>("DFMA x8", 8, out, sms, clk);
    run<3>("DMUL x4 + DADD x4", 8, out, sms, clk);
    run<4>("I2F.U32 x8 (+FADD,IADD,LOP)", 8, out, sms, clk);
    run<5>("PRMT x8", 8, out, sms, clk);
    run<6>("DP4A x8", 8, out, sms, clk);
    run<7>("IMAD x8", 8, out, sms, clk);
    run<16>("HFMA2 x8 (instr)", 8, out, sms, clk);
    run<8>("LDS.128 x8 (+4 FADD each)", 8, out, sms, clk);
    run<9>("MUFU.RCP x8", 8, out, sms, clk);
    run<14>("SHFL x8", 8, out, sms, clk);
    run<15>("I2D x4 + DADD x4", 4, out, sms, clk);
    run<10>("FFMA2 x8 + I2F x4 [fma-ops]", 16, out, sms, clk);
    run<11>("FFMA2 x8 + DFMA x4 [fma32-ops]", 16, out, sms, clk);
    run<12>("FFMA x16 + LDS.128 x2 [fma]", 16, out, sms, clk);
    run<17>("DP2A x8 (+4 IADD)", 8, out, sms, clk);
    run<18>("FFMA2 x8 alone [ffma2-instr]", 8, out, sms, clk);
    run<19>("FFMA2 x8 + FFMA x8 [ffma2-instr]", 8, out, sms, clk);
    run<20>("FFMA2 x8 + DP4A x8 [ffma2-instr]", 8, out, sms, clk);
    run_filt<1>("filter pattern FFMA2", out, sms, clk, 2);
    run_filt<1>("filter pattern FFMA2", out, sms, clk, 4);
    // The scalar twin (run_filt<0>) is NOT run: only x[i][i].x changes between iterations, so ptxas hoists every FMA-chain
    // prefix that does not depend on it out of the loop (44 of the 128 counted FMAs per iteration remain in the SASS; an
    // `asm volatile` version fared no better — PTX carries no volatile), and round 1 printed 155-169 "lane-fma/clk/SM",
    // above the 128 lanes an SM has.  The packed form cannot be split, so its rows are valid; scalar vs
    // packed pipe rates are the plain "FFMA x16" / "FFMA2 x8" rows above.
    return 0;
}
