// Scratch microbenchmark (not part of the product): issue rate of packed min/max flavours on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mnmx_probe tools/mnmx_probe.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(unsigned* sink, int iters, unsigned seed)
{
    unsigned a0 = seed ^ threadIdx.x, a1 = a0 * 3u + 1u, a2 = a0 * 5u + 7u, a3 = a0 * 7u + 3u;
    unsigned b0 = ~a0, b1 = ~a1, b2 = ~a2, b3 = ~a3;
    __half2 h0 = __halves2half2(__int2half_rn(threadIdx.x & 255), __int2half_rn(threadIdx.x >> 2)), h1 = h0, h2 = h0, h3 = h0;
    __half2 g0 = __halves2half2(__int2half_rn(seed & 127), __int2half_rn(3)), g1 = g0, g2 = g0, g3 = g0;
#pragma unroll 8
    for (int i = 0; i < iters; i++) {
        if (MODE == 0 || MODE == 2) {      // VIMNMX.S16x2
            a0 = __vmins2(a0, b1); a1 = __vmaxs2(a1, b2); a2 = __vmins2(a2, b3); a3 = __vmaxs2(a3, b0);
            b0 = __vmaxs2(b0, a1); b1 = __vmins2(b1, a2); b2 = __vmaxs2(b2, a3); b3 = __vmins2(b3, a0);
        }
        if (MODE == 1 || MODE == 2) {      // HMNMX2
            h0 = __hmin2(h0, g1); h1 = __hmax2(h1, g2); h2 = __hmin2(h2, g3); h3 = __hmax2(h3, g0);
            g0 = __hmax2(g0, h1); g1 = __hmin2(g1, h2); g2 = __hmax2(g2, h3); g3 = __hmin2(g3, h0);
        }
        if (MODE == 3) {                   // 32-bit VIMNMX (3-input when fused)
            a0 = min(a0, b1); a1 = max(a1, b2); a2 = min(a2, b3); a3 = max(a3, b0);
            b0 = max(b0, a1); b1 = min(b1, a2); b2 = max(b2, a3); b3 = min(b3, a0);
        }
        if (MODE == 4) {                   // IADD via IMAD pipe candidates
            a0 = a0 * 3u + b1; a1 = a1 * 5u + b2; a2 = a2 * 7u + b3; a3 = a3 * 9u + b0;
            b0 = b0 * 3u + a1; b1 = b1 * 5u + a2; b2 = b2 * 7u + a3; b3 = b3 * 9u + a0;
        }
    }
    unsigned r = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3 ^ *(unsigned*)&h0 ^ *(unsigned*)&h1 ^ *(unsigned*)&h2 ^ *(unsigned*)&h3 ^
                 *(unsigned*)&g0 ^ *(unsigned*)&g1 ^ *(unsigned*)&g2 ^ *(unsigned*)&g3;
    if (r == 0x12345678u) sink[0] = r;
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
    run<0>("VIMNMX.S16x2", 8);
    run<1>("HMNMX2", 8);
    run<2>("VIMNMX.S16x2 + HMNMX2", 16);
    run<3>("VIMNMX 32-bit", 8);
    run<4>("IMAD", 8);
    return 0;
}
