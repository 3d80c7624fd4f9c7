// orb_pyramid.cu -- K1: one pyramid level from the previous one, bit-exact INTER_LINEAR_EXACT (8.8 fixed point).
//
// Stage (ii) of OrbFeatureDetector::detect as called at reference src/FeatureExtractor.cpp:17 (OpenCV orb.cpp builds
// level l by resizing level l-1).  Per axis the host precomputes, in double like OpenCV, the source
// offset and the weight of the second tap (0..256) for every destination column/row; the kernel is then
//   t[r][d]  = (256-c1x[d]) * S[r][ofs_x[d]] + c1x[d] * S[r][min(ofs_x[d]+1, sw-1)]          (<= 65280)
//   D[y][d]  = ((256-c1y[y]) * t[oy][d] + c1y[y] * t[min(oy+1, sh-1)][d] + 32768) >> 16
//
// A CTA produces a 128x16 destination tile: the source rectangle it needs is staged in shared memory with 16-byte
// loads (rows are 128-byte pitched, so the vectors are always aligned and inside the row), the horizontal pass is
// kept as u16 in shared memory and the vertical pass writes 4 pixels per 32-bit store.  HBM/L2-bound: reads each
// source byte once per tile (+halo), writes each destination byte once.
#include "common.cuh"

namespace orbx {
namespace {

constexpr int PD_TW = 128;
constexpr int PD_TH = 16;
constexpr int PD_THREADS = 256;

__global__ void __launch_bounds__(PD_THREADS)
k_pyr_down(uint8_t* __restrict__ slots, size_t slot_stride, size_t src_off, int spitch, int sw, int sh, size_t dst_off,
           int dpitch, int dw, int dh, const int* __restrict__ ofs_x, const uint16_t* __restrict__ c1x,
           const int* __restrict__ ofs_y, const uint16_t* __restrict__ c1y, int s_w, int s_h)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* s_src = smem;                                         // [s_h][s_w]
    uint16_t* s_hor = (uint16_t*)(smem + (size_t)s_h * s_w);       // [s_h][PD_TW]

    const uint8_t* src = slots + blockIdx.z * slot_stride + src_off;
    uint8_t* dst = slots + blockIdx.z * slot_stride + dst_off;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * PD_TW, y0 = blockIdx.y * PD_TH;
    const int x1 = min(x0 + PD_TW, dw), y1 = min(y0 + PD_TH, dh);

    const int r0 = ofs_y[y0];
    const int r1 = min(ofs_y[y1 - 1] + 1, sh - 1);
    const int nrows = r1 - r0 + 1;
    const int c0 = ofs_x[x0] & ~15;
    const int c1 = min(ofs_x[x1 - 1] + 1, sw - 1);
    const int nvec = (c1 - c0) / 16 + 1;

    for (int i = tid; i < nrows * nvec; i += PD_THREADS) {
        int r = i / nvec, v = i - r * nvec;
        uint4 val = *reinterpret_cast<const uint4*>(src + (size_t)(r0 + r) * spitch + c0 + v * 16);
        *reinterpret_cast<uint4*>(s_src + r * s_w + v * 16) = val;
    }
    __syncthreads();

    {   // horizontal pass: a thread owns one destination column and walks the staged rows
        const int d = tid & (PD_TW - 1);
        const int x = x0 + d;
        if (x < dw) {
            const int o = ofs_x[x];
            const int w1 = c1x[x], w0 = 256 - w1;
            const int a = o - c0, b = min(o + 1, sw - 1) - c0;
            for (int r = tid / PD_TW; r < nrows; r += PD_THREADS / PD_TW)
                s_hor[r * PD_TW + d] = (uint16_t)(w0 * s_src[r * s_w + a] + w1 * s_src[r * s_w + b]);
        }
    }
    __syncthreads();

    {   // vertical pass: 4 adjacent pixels per thread and row
        const int tx = tid & 31, ty = tid >> 5;
        const int x = x0 + 4 * tx;
        if (x < dw) {
            for (int yy = ty; yy < PD_TH; yy += PD_THREADS / 32) {
                const int y = y0 + yy;
                if (y >= dh) break;
                const int oy = ofs_y[y];
                const uint32_t w1 = c1y[y], w0 = 256 - w1;
                const uint16_t* ra = s_hor + (oy - r0) * PD_TW + 4 * tx;
                const uint16_t* rb = s_hor + (min(oy + 1, sh - 1) - r0) * PD_TW + 4 * tx;
                uint32_t packed = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t v = (w0 * ra[j] + w1 * rb[j] + 32768u) >> 16;
                    packed |= v << (8 * j);
                }
                // the row pitch is a multiple of 128, so the 4-byte store stays inside the row even past dw
                *reinterpret_cast<uint32_t*>(dst + (size_t)y * dpitch + x) = packed;
            }
        }
    }
}

// Level 0 of every frame slot from a batch of device-resident frames, in ONE launch (frames are frame_pitch bytes apart,
// rows `stride` bytes apart).  16-byte vectors when the source allows it, bytes otherwise; the slot's padding columns
// [w, pitch) are never written (they stay zero from orbx_create).
__global__ void __launch_bounds__(256)
k_ingest(const uint8_t* __restrict__ src, size_t frame_pitch, size_t stride, int w, int h, uint8_t* __restrict__ slots,
         size_t slot_stride, size_t dst_off, int dpitch, int vec_ok)
{
    const uint8_t* s = src + blockIdx.z * frame_pitch;
    uint8_t* d = slots + blockIdx.z * slot_stride + dst_off;
    const int nv = vec_ok ? w >> 4 : 0;
    for (int y = blockIdx.y; y < h; y += gridDim.y) {
        const uint8_t* sr = s + (size_t)y * stride;
        uint8_t* dr = d + (size_t)y * dpitch;
        for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x)
            reinterpret_cast<uint4*>(dr)[v] = __ldcs(reinterpret_cast<const uint4*>(sr) + v);
        for (int x = nv * 16 + blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x) dr[x] = sr[x];
    }
}

}  // namespace

cudaError_t launch_ingest(const uint8_t* d_frames, size_t frame_pitch, size_t stride, int w, int h, uint8_t* slots,
                          size_t slot_stride, const LevelGeom& L0, int nframes, cudaStream_t s)
{
    const int vec_ok = (((uintptr_t)d_frames | frame_pitch | stride) & 15) == 0;
    dim3 grid(div_up(div_up(w, 16), 256), h < 270 ? h : 270, nframes);
    k_ingest<<<grid, 256, 0, s>>>(d_frames, frame_pitch, stride, w, h, slots, slot_stride, L0.img_off, L0.pitch, vec_ok);
    return cudaGetLastError();
}

// Largest staged source rectangle over all tiles of a level (host tables), so shared memory can be sized exactly.
void pyr_down_smem_extent(const int* ofs_x, int dw, const int* ofs_y, int dh, int sw, int sh, int* s_w, int* s_h)
{
    int mw = 0, mh = 0;
    for (int x0 = 0; x0 < dw; x0 += PD_TW) {
        int x1 = x0 + PD_TW < dw ? x0 + PD_TW : dw;
        int c0 = ofs_x[x0] & ~15;
        int c1 = ofs_x[x1 - 1] + 1 < sw - 1 ? ofs_x[x1 - 1] + 1 : sw - 1;
        int wv = ((c1 - c0) / 16 + 1) * 16;
        if (wv > mw) mw = wv;
    }
    for (int y0 = 0; y0 < dh; y0 += PD_TH) {
        int y1 = y0 + PD_TH < dh ? y0 + PD_TH : dh;
        int r0 = ofs_y[y0];
        int r1 = ofs_y[y1 - 1] + 1 < sh - 1 ? ofs_y[y1 - 1] + 1 : sh - 1;
        if (r1 - r0 + 1 > mh) mh = r1 - r0 + 1;
    }
    *s_w = mw;
    *s_h = mh;
}

cudaError_t launch_pyr_down_level(uint8_t* slots, size_t slot_stride, const LevelGeom& src, const LevelGeom& dst, int s_w,
                                  int s_h, int nframes, cudaStream_t s)
{
    dim3 grid(div_up(dst.w, PD_TW), div_up(dst.h, PD_TH), nframes);
    size_t smem = (size_t)s_h * s_w + (size_t)s_h * PD_TW * sizeof(uint16_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_pyr_down, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_pyr_down<<<grid, PD_THREADS, smem, s>>>(slots, slot_stride, src.img_off, src.pitch, src.w, src.h, dst.img_off, dst.pitch,
                                               dst.w, dst.h, dst.ofs_x, dst.c1x, dst.ofs_y, dst.c1y, s_w, s_h);
    return cudaGetLastError();
}

}  // namespace orbx
