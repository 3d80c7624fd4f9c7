// Scratch microbenchmark (not part of the product): issue rate of packed min/max flavours on sm_100a, and whether
// HMNMX2 orders 16-bit integer lanes (fp16 denormal bit patterns) like VIMNMX.U16x2 does.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mnmx_probe tools/mnmx_probe.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hmin2u(unsigned a, unsigned b)
{
    unsigned d;
    asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ unsigned hmax2u(unsigned a, unsigned b)
{
    unsigned d;
    asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(1024) k(unsigned* sink, int iters, unsigned seed)
{
    unsigned a0 = (seed ^ threadIdx.x) & 0x00FF00FFu, a1 = (a0 * 3u + 1u) & 0x00FF00FFu, a2 = (a0 * 5u + 7u) & 0x00FF00FFu,
             a3 = (a0 * 7u + 3u) & 0x00FF00FFu;
    unsigned b0 = ~a0 & 0x00FF00FFu, b1 = ~a1 & 0x00FF00FFu, b2 = ~a2 & 0x00FF00FFu, b3 = ~a3 & 0x00FF00FFu;
    unsigned h0 = a0 ^ 0x00110022u, h1 = a1 ^ 0x00110022u, h2 = a2 ^ 0x00330044u, h3 = a3 ^ 0x00050006u;
    unsigned g0 = b0 ^ 0x00110022u, g1 = b1 ^ 0x00110022u, g2 = b2 ^ 0x00330044u, g3 = b3 ^ 0x00050006u;
#pragma unroll 8
    for (int i = 0; i < iters; i++) {
        if (MODE == 0 || MODE == 2) {      // VIMNMX.S16x2
            a0 = __vmins2(a0, b1); a1 = __vmaxs2(a1, b2); a2 = __vmins2(a2, b3); a3 = __vmaxs2(a3, b0);
            b0 = __vmaxs2(b0, a1); b1 = __vmins2(b1, a2); b2 = __vmaxs2(b2, a3); b3 = __vmins2(b3, a0);
        }
        if (MODE == 1 || MODE == 2 || MODE == 6) {      // HMNMX2 on integer lanes
            h0 = hmin2u(h0, g1); h1 = hmax2u(h1, g2); h2 = hmin2u(h2, g3); h3 = hmax2u(h3, g0);
            g0 = hmax2u(g0, h1); g1 = hmin2u(g1, h2); g2 = hmax2u(g2, h3); g3 = hmin2u(g3, h0);
        }
        if (MODE == 3) {                   // 32-bit VIMNMX
            a0 = min(a0, b1); a1 = max(a1, b2); a2 = min(a2, b3); a3 = max(a3, b0);
            b0 = max(b0, a1); b1 = min(b1, a2); b2 = max(b2, a3); b3 = min(b3, a0);
        }
        if (MODE == 4) {                   // IMAD
            a0 = a0 * 3u + b1; a1 = a1 * 5u + b2; a2 = a2 * 7u + b3; a3 = a3 * 9u + b0;
            b0 = b0 * 3u + a1; b1 = b1 * 5u + a2; b2 = b2 * 7u + a3; b3 = b3 * 9u + a0;
        }
        if (MODE == 5 || MODE == 6) {      // VIMNMX3.S16x2
            a0 = __vimin3_s16x2(a0, b1, b2); a1 = __vimax3_s16x2(a1, b2, b3); a2 = __vimin3_s16x2(a2, b3, b0); a3 = __vimax3_s16x2(a3, b0, b1);
            b0 = __vimax3_s16x2(b0, a1, a2); b1 = __vimin3_s16x2(b1, a2, a3); b2 = __vimax3_s16x2(b2, a3, a0); b3 = __vimin3_s16x2(b3, a0, a1);
        }
        if (MODE == 7) {                   // PRMT
            a0 = __byte_perm(a0, b1, 0x5432); a1 = __byte_perm(a1, b2, 0x4140); a2 = __byte_perm(a2, b3, 0x5432); a3 = __byte_perm(a3, b0, 0x1054);
            b0 = __byte_perm(b0, a1, 0x5432); b1 = __byte_perm(b1, a2, 0x4140); b2 = __byte_perm(b2, a3, 0x5432); b3 = __byte_perm(b3, a0, 0x1054);
        }
    }
    unsigned r = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3 ^ h0 ^ h1 ^ h2 ^ h3 ^ g0 ^ g1 ^ g2 ^ g3;
    if (r == 0x12345678u) sink[0] = r;
}

// exhaustive check over all 16-bit pairs in [0, 0x3FF]: min.f16x2 / max.f16x2 on the bit patterns == unsigned min/max
__global__ void k_check(unsigned* bad)
{
    unsigned a = blockIdx.x * blockDim.x + threadIdx.x;   // 0 .. 1023
    if (a >= 1024) return;
    unsigned n = 0;
    for (unsigned b = 0; b < 1024; b++) {
        unsigned x = a | (b << 16), y = b | (a << 16);
        unsigned mn = hmin2u(x, y), mx = hmax2u(x, y);
        unsigned emn = min(a, b) | (min(a, b) << 16), emx = max(a, b) | (max(a, b) << 16);
        n += (mn != emn) + (mx != emx);
    }
    if (n) atomicAdd(bad, n);
}

template <int MODE>
static void run(const char* name, int ops_per_iter)
{
    unsigned* sink;
    cudaMalloc(&sink, 256);
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 15, blocks = sms * 2, threads = 1024;
    k<MODE><<<blocks, threads>>>(sink, 64, 1);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(sink, iters, 7 + r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double winst = (double)blocks * threads / 32 * iters * ops_per_iter;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %.3f ms  %.1f G warp-inst/s  = %.2f warp-inst/clk/SM @%.0f MHz (nominal)\n", name, best, winst / best / 1e6,
           winst / (best * 1e-3) / sms / (clk * 1e3), clk / 1e3);
    cudaFree(sink);
}

int main()
{
    unsigned* bad;
    cudaMalloc(&bad, 4);
    cudaMemset(bad, 0, 4);
    k_check<<<4, 256>>>(bad);
    unsigned hbad = 1;
    cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
    printf("min/max.f16x2 on integer lanes 0..1023: %u mismatches vs unsigned min/max\n", hbad);
    run<0>("VIMNMX.S16x2", 8);
    run<1>("HMNMX2", 8);
    run<2>("VIMNMX.S16x2 + HMNMX2", 16);
    run<3>("VIMNMX 32-bit", 8);
    run<4>("IMAD", 8);
    run<5>("VIMNMX3.S16x2", 8);
    run<6>("VIMNMX3.S16x2 + HMNMX2", 16);
    run<7>("PRMT", 8);
    return 0;
}
