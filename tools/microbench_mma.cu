// Warp-level tensor-pipe microbenchmarks for sm_100a (B200): the rate of the legacy mma.sync path
// (HMMA / TF32), of the fp32 -> packed-fp16 conversion that feeds it, and how both co-issue with the FP32 FMA
// pipe.  Decides whether the separable SSIM / blur filters can move off the FMA pipe as banded Toeplitz
// contractions (DESIGN.md "tensor-core question").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_mma microbench_mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_f16_k8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(b[0]));
}

// MODE 0: HMMA.16816.F32 (f16) x8 independent accumulators; 1: bf16; 2: TF32 m16n8k8; 3: f16 m16n8k8
// MODE 4: cvt.rn.f16x2.f32 x8 (F2FP pack); 5: LOP3 x8; 6: HMMA x8 + FFMA2 x16 co-issue; 7: HMMA x8 + F2FP x8 + LOP x8 + FADD x8
// MODE 8: HMMA x4 dependent pairs (latency chain)
template <int MODE>
__global__ void __launch_bounds__(256) bench(float *out, int iters, float seed) {
    float d[8][4];
    uint32_t a[4], b[2];
    float f[16];
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) d[i][j] = seed * i + j;
        u[i] = threadIdx.x * 2654435761u + i;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) f[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = 0x3c003c00u + (threadIdx.x & 1);
    b[0] = 0x38003800u; b[1] = 0x34003400u;
    const float w = seed * 0.5f, c = seed * 0.25f;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 6 || MODE == 7) {
#pragma unroll
            for (int i = 0; i < 8; i++) mma_f16(d[i], a, b);
        }
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) mma_bf16(d[i], a, b);
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) mma_tf32(d[i], a, b);
        }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; i++) mma_f16_k8(d[i], a, b);
        }
        if (MODE == 8) {
#pragma unroll
            for (int i = 0; i < 4; i++) { mma_f16(d[0], a, b); mma_f16(d[1], a, b); }
        }
        if (MODE == 4 || MODE == 7) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t h;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(f[2 * i]), "f"(f[2 * i + 1]));
                u[i] ^= h;
                if (MODE == 4) f[2 * i] += 1.0f;
            }
        }
        if (MODE == 5 || MODE == 7) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(0x0f0f0f0f));
        }
        if (MODE == 7) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(w));
        }
        if (MODE == 6) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned long long p, q, r;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(f[2 * i]), "f"(f[2 * i + 1]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(q) : "f"(w));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(q), "l"(r));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(q), "l"(r));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(f[2 * i]), "=f"(f[2 * i + 1]) : "l"(p));
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += d[i][0] + d[i][1] + d[i][2] + d[i][3] + (float)u[i];
#pragma unroll
    for (int i = 0; i < 16; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ldmatrix.x4 rate (shared-memory fragments)
__global__ void __launch_bounds__(256) bench_ldsm(float *out, int iters) {
    __shared__ __align__(16) uint16_t sm[8 * 1024];
    for (int i = threadIdx.x; i < 8 * 1024; i += 256) sm[i] = (uint16_t)i;
    __syncthreads();
    uint32_t acc = 0;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 512;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t r0, r1, r2, r3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + ((i * 4096 + it * 16) & 8191)));
            acc += r0 ^ r1 ^ r2 ^ r3;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
}

template <int MODE>
static int run(const char *name, double unitsPerIter, const char *unit, float *out, int sms, int clk_khz, int blocksPerSM) {
    int blocks = sms * blocksPerSM;
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
    // warp-instructions per clock per SM
    double warpInstr = unitsPerIter * ITERS * 8.0 * blocks;
    double perClkSM = warpInstr / (best * 1e-3) / sms / (clk_khz * 1e3);
    printf("%-44s %2d warps/SM %8.3f ms  %7.4f %s/clk/SM  (1 per %.2f clk per SMSP)\n", name, blocksPerSM * 8, best, perClkSM, unit,
           4.0 / perClkSM);
    return 0;
}

int main() {
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s SMs=%d clk=%d kHz (rates quoted at max clock)\n", p.name, p.multiProcessorCount, clk);
    float *out;
    CHECK(cudaMalloc(&out, sizeof(float) * 256 * p.multiProcessorCount * 8));
    int sms = p.multiProcessorCount;
    for (int bps = 1; bps <= 4; bps *= 2) {
        run<0>("HMMA m16n8k16 f16->f32 x8 [warp-mma]", 8, "mma", out, sms, clk, bps);
    }
    run<1>("HMMA m16n8k16 bf16->f32 x8 [warp-mma]", 8, "mma", out, sms, clk, 2);
    run<2>("HMMA m16n8k8 tf32->f32 x8 [warp-mma]", 8, "mma", out, sms, clk, 2);
    run<3>("HMMA m16n8k8 f16->f32 x8 [warp-mma]", 8, "mma", out, sms, clk, 2);
    run<8>("HMMA m16n8k16 f16 2 chains x4 [warp-mma]", 8, "mma", out, sms, clk, 1);
    run<4>("F2FP cvt.rn.f16x2.f32 x8 (+LOP,FADD) [instr]", 8, "instr", out, sms, clk, 2);
    run<5>("LOP3 x8 [instr]", 8, "instr", out, sms, clk, 2);
    run<6>("HMMA x8 + FFMA2 x16 [warp-mma]", 8, "mma", out, sms, clk, 2);
    run<7>("HMMA x8 + F2FP x8 + LOP x16 + FADD x8 [mma]", 8, "mma", out, sms, clk, 2);
    {
        int blocks = sms * 2;
        bench_ldsm<<<blocks, 256>>>(out, 64);
        CHECK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int r = 0; r < 5; r++) {
            cudaEventRecord(e0); bench_ldsm<<<blocks, 256>>>(out, ITERS); cudaEventRecord(e1);
            CHECK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double bytes = 8.0 * 512 * ITERS * 8.0 * blocks;
        printf("%-44s %8.3f ms  %7.2f B/clk/SM\n", "LDSM.x4 x8 (512 B per warp-instr)", best, bytes / (best * 1e-3) / sms / (clk * 1e3));
    }
    return 0;
}
