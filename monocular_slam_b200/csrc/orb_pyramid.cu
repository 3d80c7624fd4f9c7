// orb_pyramid.cu -- K1: one pyramid level from the previous one, bit-exact INTER_LINEAR_EXACT (8.8 fixed point).
//
// Stage (ii) of OrbFeatureDetector::detect as called at reference src/FeatureExtractor.cpp:17 (OpenCV orb.cpp builds
// level l by resizing level l-1).  Per axis the host precomputes, in double like OpenCV, the source
// offset and the weight of the second tap (0..256) for every destination column/row.  OpenCV interpolates rows first:
//   t[r][d]  = (256-c1x[d]) * S[r][ofs_x[d]] + c1x[d] * S[r][min(ofs_x[d]+1, sw-1)]          (<= 65280, exact)
//   D[y][d]  = ((256-c1y[y]) * t[oy][d] + c1y[y] * t[min(oy+1, sh-1)][d] + 32768) >> 16
// Everything before the final shift is exact integer arithmetic, so the two passes commute; this kernel runs the
// VERTICAL pass first because it vectorises: a thread takes 16 source pixels of the two source rows of a destination
// row straight from global memory (two 16-byte loads), widens them to 16x2 lanes and forms
// V = w0y * a + w1y * b with two 32-bit IMADs per lane pair (each lane stays <= 65280, so lanes never carry into each
// other), and stores V as u16 in shared memory.  The horizontal pass then produces 2 adjacent destination pixels per
// thread: two u16 loads and two IMADs per pixel, the result byte is byte 2 of the sum, and one PRMT packs the
// 16-bit store.  About 10 instructions per destination pixel; HBM/L2-bound (reads the source once from HBM, ~1.6x
// from L1/L2; writes each destination byte once).
#include "common.cuh"

namespace orbx {
namespace {

constexpr int PD_TW = 128;
constexpr int PD_TH = 64;
constexpr int PD_THREADS = 256;

__device__ __forceinline__ uint32_t lanes_lo(uint32_t g) { return __byte_perm(g, 0u, 0x4140); }
__device__ __forceinline__ uint32_t lanes_hi(uint32_t g) { return __byte_perm(g, 0u, 0x4342); }

// PITCH: row pitch of the intermediate in u16 entries, a compile-time constant so that the horizontal pass addresses its
// rows with immediate offsets (192 covers scale factors up to ~1.37, 320 up to 2).
template <int PITCH>
__global__ void __launch_bounds__(PD_THREADS)
k_pyr_down(uint8_t* __restrict__ slots, size_t slot_stride, size_t src_off, int spitch, int sw, int sh, size_t dst_off,
           int dpitch, int dw, int dh, const int* __restrict__ ofs_x, const uint16_t* __restrict__ c1x,
           const int* __restrict__ ofs_y, const uint16_t* __restrict__ c1y)
{
    __shared__ __align__(16) uint16_t s_v[PD_TH * PITCH];

    const uint8_t* src = slots + blockIdx.z * slot_stride + src_off;
    uint8_t* dst = slots + blockIdx.z * slot_stride + dst_off;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * PD_TW, y0 = blockIdx.y * PD_TH;
    const int x1 = min(x0 + PD_TW, dw), rows = min(PD_TH, dh - y0);
    const int c0 = __ldg(ofs_x + x0) & ~15;                        // first staged source column (16-byte aligned)
    const int c1 = min(__ldg(ofs_x + x1 - 1) + 1, sw - 1);         // last source column any tap reads
    const int nvec = (c1 - c0) / 16 + 1;                           // <= PITCH / 16 (checked by the launcher)

    // this thread's 2 destination columns (horizontal pass); loaded early so the latency overlaps the vertical pass.
    // Two pixels per lane keep a warp's u16 gathers within ~40 shared-memory words (1-2 wavefronts per load); with four
    // per lane they spread over ~80 words and the kernel was bound by L1 wavefronts.
    const int tx = tid & 63, ty = tid >> 6;
    const int x = x0 + 2 * tx;
    int o[2];
    uint32_t wx1[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int xi = min(x + j, dw - 1);
        o[j] = __ldg(ofs_x + xi) - c0;
        wx1[j] = __ldg(c1x + xi);
    }

    // ---- vertical pass: V[y][c] = w0y * S[oy][c] + w1y * S[oy+1][c]; one item = 16 source pixels of one destination row.
    // A thread owns up to ITEMS items; their table lookups, then all their 16-byte pixel loads, are issued back to back
    // before anything is consumed (item after item would serialise two dependent global-load latencies per item).
    {
        constexpr int ITEMS = (PD_TH * (PITCH / 16) + PD_THREADS - 1) / PD_THREADS;
        const uint32_t rcp = (65536u + nvec - 1) / nvec;           // item / nvec == (item * rcp) >> 16 for item < 64 * nvec, nvec <= 36 (checked exhaustively)
        const int nitems = rows * nvec;
        int sidx[ITEMS];                                           // smem index of the item's first entry, -1: no item
        uint32_t w1[ITEMS];
        const uint8_t* pa[ITEMS];
        const uint8_t* pb[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const int item = tid + i * PD_THREADS;
            const bool ok = item < nitems;
            const int yy = ok ? (int)(((uint32_t)item * rcp) >> 16) : 0, v = ok ? item - yy * nvec : 0;
            const int oy = __ldg(ofs_y + y0 + yy);
            w1[i] = __ldg(c1y + y0 + yy);
            pa[i] = src + (size_t)oy * spitch + c0 + v * 16;
            pb[i] = src + (size_t)min(oy + 1, sh - 1) * spitch + c0 + v * 16;
            sidx[i] = ok ? yy * PITCH + v * 16 : -1;
        }
        uint4 a[ITEMS], b[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            if (sidx[i] >= 0) {
                a[i] = *reinterpret_cast<const uint4*>(pa[i]);
                b[i] = *reinterpret_cast<const uint4*>(pb[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            if (sidx[i] >= 0) {
                const uint32_t w0 = 256u - w1[i], ww = w1[i];
                uint4 lo, hi;
                lo.x = lanes_lo(a[i].x) * w0 + lanes_lo(b[i].x) * ww; lo.y = lanes_hi(a[i].x) * w0 + lanes_hi(b[i].x) * ww;
                lo.z = lanes_lo(a[i].y) * w0 + lanes_lo(b[i].y) * ww; lo.w = lanes_hi(a[i].y) * w0 + lanes_hi(b[i].y) * ww;
                hi.x = lanes_lo(a[i].z) * w0 + lanes_lo(b[i].z) * ww; hi.y = lanes_hi(a[i].z) * w0 + lanes_hi(b[i].z) * ww;
                hi.z = lanes_lo(a[i].w) * w0 + lanes_lo(b[i].w) * ww; hi.w = lanes_hi(a[i].w) * w0 + lanes_hi(b[i].w) * ww;
                uint4* out = reinterpret_cast<uint4*>(s_v + sidx[i]);
                out[0] = lo;
                out[1] = hi;
            }
        }
    }
    __syncthreads();

    // ---- horizontal pass: 2 adjacent destination pixels per thread, rows ty, ty + 4, ...
    if (x < dw) {
        const uint32_t wa0 = 256u - wx1[0], wa1 = 256u - wx1[1];
        const uint16_t* p0 = s_v + ty * PITCH + o[0];
        const uint16_t* p1 = s_v + ty * PITCH + o[1];
        uint8_t* drow = dst + (size_t)(y0 + ty) * dpitch + x;
#pragma unroll
        for (int k = 0; k < PD_TH / 4; k++) {
            if (ty + 4 * k < rows) {
                // the second tap of a clamped column has weight 0 (it may read one entry of slack)
                const uint32_t t0 = (uint32_t)p0[4 * k * PITCH] * wa0 + ((uint32_t)p0[4 * k * PITCH + 1] * wx1[0] + 32768u);
                const uint32_t t1 = (uint32_t)p1[4 * k * PITCH] * wa1 + ((uint32_t)p1[4 * k * PITCH + 1] * wx1[1] + 32768u);
                // result byte = bits 16..23 of each sum; the row pitch is a multiple of 128, so the 2-byte store stays inside
                // the row even past dw
                *reinterpret_cast<uint16_t*>(drow + (size_t)(4 * k) * dpitch) = (uint16_t)__byte_perm(t0, t1, 0x0062);
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

// Same for 3-channel frames (cv::Mat CV_8UC3, BGR interleaved): cvtColor(BGR2GRAY) as cv::ORB applies it to non-gray
// input, OpenCV's 15-bit fixed point gray = (B*3735 + G*19235 + R*9798 + 2^14) >> 15.  4 pixels (12 bytes) per thread.
__global__ void __launch_bounds__(256)
k_ingest_bgr(const uint8_t* __restrict__ src, size_t frame_pitch, size_t stride, int w, int h, uint8_t* __restrict__ slots,
             size_t slot_stride, size_t dst_off, int dpitch, int word_ok)
{
    const uint8_t* s = src + blockIdx.z * frame_pitch;
    uint8_t* d = slots + blockIdx.z * slot_stride + dst_off;
    const int ngroups = (w + 3) >> 2;
    for (int y = blockIdx.y; y < h; y += gridDim.y) {
        const uint8_t* sr = s + (size_t)y * stride;
        uint8_t* dr = d + (size_t)y * dpitch;
        for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += gridDim.x * blockDim.x) {
            const int x = 4 * g;
            uint32_t px[12];
            if (word_ok && x + 4 <= w) {
                const uint32_t* p = reinterpret_cast<const uint32_t*>(sr + 3 * x);
                const uint32_t a = __ldcs(p), b = __ldcs(p + 1), c = __ldcs(p + 2);
#pragma unroll
                for (int i = 0; i < 4; i++) { px[i] = (a >> (8 * i)) & 255u; px[4 + i] = (b >> (8 * i)) & 255u; px[8 + i] = (c >> (8 * i)) & 255u; }
            } else {
#pragma unroll
                for (int i = 0; i < 12; i++) px[i] = (x + i / 3 < w) ? sr[3 * x + i] : 0u;
            }
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; j++)
                packed |= ((px[3 * j] * 3735u + px[3 * j + 1] * 19235u + px[3 * j + 2] * 9798u + (1u << 14)) >> 15) << (8 * j);
            if (x + 4 <= w) *reinterpret_cast<uint32_t*>(dr + x) = packed;      // rows are 128-byte pitched: aligned
            else for (int j = 0; x + j < w; j++) dr[x + j] = (uint8_t)(packed >> (8 * j));   // keep the padding columns zero
        }
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

cudaError_t launch_ingest_bgr(const uint8_t* d_frames, size_t frame_pitch, size_t stride, int w, int h, uint8_t* slots,
                              size_t slot_stride, const LevelGeom& L0, int nframes, cudaStream_t s)
{
    const int word_ok = (((uintptr_t)d_frames | frame_pitch | stride) & 3) == 0;
    dim3 grid(div_up(div_up(w, 4), 256), h < 270 ? h : 270, nframes);
    k_ingest_bgr<<<grid, 256, 0, s>>>(d_frames, frame_pitch, stride, w, h, slots, slot_stride, L0.img_off, L0.pitch, word_ok);
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
    (void)s_h;
    dim3 grid(div_up(dst.w, PD_TW), div_up(dst.h, PD_TH), nframes);
    if (s_w + 8 <= 192)
        k_pyr_down<192><<<grid, PD_THREADS, 0, s>>>(slots, slot_stride, src.img_off, src.pitch, src.w, src.h, dst.img_off, dst.pitch, dst.w,
                                                     dst.h, dst.ofs_x, dst.c1x, dst.ofs_y, dst.c1y);
    else if (s_w + 8 <= 320)
        k_pyr_down<320><<<grid, PD_THREADS, 0, s>>>(slots, slot_stride, src.img_off, src.pitch, src.w, src.h, dst.img_off, dst.pitch, dst.w,
                                                     dst.h, dst.ofs_x, dst.c1x, dst.ofs_y, dst.c1y);
    else
        return cudaErrorInvalidValue;     // scale factors above 2 are rejected by orbx_create
    return cudaGetLastError();
}

}  // namespace orbx
