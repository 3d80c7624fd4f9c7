// orb_fast.cu -- K2: FAST-9/16 detection, integer corner score, 3x3 strict non-max suppression, border filter,
// compaction -- every pyramid level of every frame of the batch in ONE launch.
//
// Stages (iii)-(iv) of OrbFeatureDetector::detect (reference src/FeatureExtractor.cpp:17; OpenCV fast.cpp /
// fast_score.cpp):
//   d[k] = I(p) - I(ring k);  m = max over the 16 arcs of 9 contiguous ring pixels of max(min d, min -d)
//   corner  <=>  m > threshold (20);  score = m - 1;  keep iff score > all 8 neighbours' scores (non-corners score 0)
//   and 31 <= x < w-31, 31 <= y < h-31.
//
// A CTA owns a 128x32 tile of the interior [31,w-31) x [31,h-31) of one level.  It stages the tile plus a 4-pixel
// halo in shared memory with 16-byte loads (rows are 128-byte pitched), then
//   1a. every thread tests ~17 pixels of the (tile + 1) score region: the 16 ring comparisons are folded into two
//       16-bit masks and the 9-contiguous test is 4 shift-ANDs on the doubled mask; corners are appended to a
//       shared-memory queue with one ballot-aggregated atomic per warp,
//   1b. the queue is drained densely (no divergence): sliding-window minimum over the ring gives m exactly,
//   2.  NMS on the score tile; survivors are staged in shared memory, the CTA reserves its range of the level's
//       candidate list with ONE global atomic and writes (x, y, score); the per-level score histogram used by the
//       retainBest cut (K3) is accumulated with one RED per survivor.
// Bound: integer ALU / shared-memory issue, not HBM (each level byte is read 1.3x from L2, once from HBM).
#include "common.cuh"

namespace orbx {
namespace {

constexpr int FT_TW = 128;
constexpr int FT_TH = 32;
constexpr int FT_THREADS = 256;
constexpr int FT_SP = 160;               // smem image pitch: 11 (alignment) + 4 + 128 + 4 = 147 -> 10 vectors
constexpr int FT_SH = FT_TH + 8;         // 40 rows
constexpr int FT_CW = FT_TW + 2;         // score region 130 x 34
constexpr int FT_CH = FT_TH + 2;
constexpr int FT_CP = 132;               // score pitch
constexpr int FT_NPOS = FT_CW * FT_CH;   // 4420
constexpr int FT_EMIT = (FT_TW / 2 + 1) * (FT_TH / 2 + 1);   // NMS allows at most one maximum per 2x2 block

__device__ __forceinline__ bool arc9(uint32_t m)
{
    m |= m << 16;
    uint32_t r = m & (m >> 1);
    r &= r >> 2;
    r &= r >> 4;          // bit i: ring pixels i..i+7 all set
    r &= m >> 8;          // ... and i+8
    return (r & 0xFFFFu) != 0;
}

__global__ void __launch_bounds__(FT_THREADS)
k_fast(const __grid_constant__ FrameGeom g, const uint8_t* __restrict__ slots, size_t slot_stride, Cand* __restrict__ cand,
       size_t cand_stride, FrameCounters* __restrict__ ctr)
{
    __shared__ __align__(16) uint8_t s_img[FT_SH * FT_SP];
    __shared__ uint8_t s_score[FT_CH * FT_CP];
    __shared__ uint16_t s_queue[FT_NPOS];
    __shared__ Cand s_emit[FT_EMIT];
    __shared__ int s_qn, s_en, s_base;

    const int tid = threadIdx.x, lane = tid & 31;
    const int frame = blockIdx.y;

    int level = 0;
    while (level + 1 < g.nlevels && (int)blockIdx.x >= g.lv[level + 1].tile_start) level++;
    const LevelGeom& L = g.lv[level];
    const int t = blockIdx.x - L.tile_start;
    const int ty = t / L.tiles_x, tx = t - ty * L.tiles_x;
    const int w = L.w, h = L.h, pitch = L.pitch;
    const uint8_t* img = slots + frame * slot_stride + L.img_off;
    const int ox = 31 + tx * FT_TW, oy = 31 + ty * FT_TH;   // first output pixel of the tile
    const int gx0 = ox - 15, gy0 = oy - 4;                  // image coords of s_img[0][0]; gx0 is a multiple of 16
    const int thr = g.fast_threshold;

    if (tid == 0) { s_qn = 0; s_en = 0; }
    for (int i = tid; i < FT_SH * (FT_SP / 16); i += FT_THREADS) {
        int r = i / (FT_SP / 16), v = i - r * (FT_SP / 16);
        int gy = gy0 + r, gx = gx0 + v * 16;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (gy < h && gx < pitch) val = *reinterpret_cast<const uint4*>(img + (size_t)gy * pitch + gx);
        *reinterpret_cast<uint4*>(s_img + r * FT_SP + v * 16) = val;
    }
    for (int i = tid; i < FT_CH * FT_CP / 4; i += FT_THREADS) reinterpret_cast<uint32_t*>(s_score)[i] = 0;
    __syncthreads();

    // ---- 1a: corner test over the score region (tile + 1 pixel each side)
    for (int p0 = 0; p0 < FT_NPOS; p0 += FT_THREADS) {
        const int p = p0 + tid;
        bool corner = false, dark = false;
        if (p < FT_NPOS) {
            const int cy = p / FT_CW, cx = p - cy * FT_CW;
            const int x = ox - 1 + cx, y = oy - 1 + cy;
            if (x <= w - 31 && y <= h - 31) {
                const uint8_t* c = s_img + (cy + 3) * FT_SP + (cx + 14);
                const int v = c[0], hi = v + thr, lo = v - thr;
                // compass pixels k = 0, 4, 8, 12: any 9-arc contains at least two of them
                const int q0 = c[3 * FT_SP], q4 = c[3], q8 = c[-3 * FT_SP], q12 = c[-3];
                const int nb = (q0 > hi) + (q4 > hi) + (q8 > hi) + (q12 > hi);
                const int nd = (q0 < lo) + (q4 < lo) + (q8 < lo) + (q12 < lo);
                if (nb >= 2 || nd >= 2) {
                    int r[16];
                    r[0] = q0; r[4] = q4; r[8] = q8; r[12] = q12;
                    r[1] = c[3 * FT_SP + 1];  r[2] = c[2 * FT_SP + 2];   r[3] = c[FT_SP + 3];
                    r[5] = c[-FT_SP + 3];     r[6] = c[-2 * FT_SP + 2];  r[7] = c[-3 * FT_SP + 1];
                    r[9] = c[-3 * FT_SP - 1]; r[10] = c[-2 * FT_SP - 2]; r[11] = c[-FT_SP - 3];
                    r[13] = c[FT_SP - 3];     r[14] = c[2 * FT_SP - 2];  r[15] = c[3 * FT_SP - 1];
                    uint32_t bm = 0, dm = 0;
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        bm |= (uint32_t)(r[k] > hi) << k;
                        dm |= (uint32_t)(r[k] < lo) << k;
                    }
                    dark = arc9(dm);
                    corner = dark || arc9(bm);
                }
            }
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, corner);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_qn, __popc(bal));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (corner) s_queue[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)(p | (dark ? 0x8000 : 0));
        }
    }
    __syncthreads();

    // ---- 1b: exact score of the queued corners
    const int qn = s_qn;
    for (int i = tid; i < qn; i += FT_THREADS) {
        const int e = s_queue[i];
        const int p = e & 0x7FFF;
        const int cy = p / FT_CW, cx = p - cy * FT_CW;
        const uint8_t* c = s_img + (cy + 3) * FT_SP + (cx + 14);
        const int v = c[0];
        int d[16];
        d[0] = c[3 * FT_SP];       d[1] = c[3 * FT_SP + 1];   d[2] = c[2 * FT_SP + 2];   d[3] = c[FT_SP + 3];
        d[4] = c[3];               d[5] = c[-FT_SP + 3];      d[6] = c[-2 * FT_SP + 2];  d[7] = c[-3 * FT_SP + 1];
        d[8] = c[-3 * FT_SP];      d[9] = c[-3 * FT_SP - 1];  d[10] = c[-2 * FT_SP - 2]; d[11] = c[-FT_SP - 3];
        d[12] = c[-3];             d[13] = c[FT_SP - 3];      d[14] = c[2 * FT_SP - 2];  d[15] = c[3 * FT_SP - 1];
        const bool dk = (e & 0x8000) != 0;
#pragma unroll
        for (int k = 0; k < 16; k++) d[k] = dk ? v - d[k] : d[k] - v;   // positive on the corner's arc
        int t3[16];
#pragma unroll
        for (int k = 0; k < 16; k++) t3[k] = min(min(d[k], d[(k + 1) & 15]), d[(k + 2) & 15]);
        int m = -256;
#pragma unroll
        for (int k = 0; k < 16; k++) m = max(m, min(min(t3[k], t3[(k + 3) & 15]), t3[(k + 6) & 15]));
        s_score[cy * FT_CP + cx] = (uint8_t)(m - 1);   // m in (thr, 255]
    }
    __syncthreads();

    // ---- 2: non-max suppression on the tile proper
    for (int p = tid; p < FT_TW * FT_TH; p += FT_THREADS) {
        const int yy = p / FT_TW, xx = p - yy * FT_TW;
        const int x = ox + xx, y = oy + yy;
        const uint8_t* sp = s_score + (yy + 1) * FT_CP + (xx + 1);
        const int s = sp[0];
        if (s == 0 || x >= w - 31 || y >= h - 31) continue;
        if (s > sp[-1] && s > sp[1] && s > sp[-FT_CP - 1] && s > sp[-FT_CP] && s > sp[-FT_CP + 1] && s > sp[FT_CP - 1] &&
            s > sp[FT_CP] && s > sp[FT_CP + 1]) {
            int pos = atomicAdd(&s_en, 1);
            Cand cnd;
            cnd.xy = ((uint32_t)y << 16) | (uint32_t)x;
            cnd.score = (uint32_t)s;
            s_emit[pos] = cnd;
        }
    }
    __syncthreads();
    const int en = s_en;
    if (en == 0) return;
    if (tid == 0) s_base = atomicAdd(&ctr[frame].ncand[level], en);
    __syncthreads();
    const int base = s_base;
    Cand* out = cand + frame * cand_stride + L.cand_off;
    for (int i = tid; i < en; i += FT_THREADS) {
        const Cand cnd = s_emit[i];
        if (base + i < L.cand_cap) out[base + i] = cnd;
        else atomicOr(&ctr[frame].overflow, 1);
        atomicAdd(&ctr[frame].hist[level][cnd.score], 1u);
    }
}

}  // namespace

cudaError_t launch_fast(const FrameGeom& g, uint8_t* slots, size_t slot_stride, Cand* cand, size_t cand_stride,
                        FrameCounters* ctr, int nframes, cudaStream_t s)
{
    if (g.total_tiles == 0) return cudaSuccess;
    dim3 grid(g.total_tiles, nframes);
    k_fast<<<grid, FT_THREADS, 0, s>>>(g, slots, slot_stride, cand, cand_stride, ctr);
    return cudaGetLastError();
}

void fast_tile_dims(int* tw, int* th) { *tw = FT_TW; *th = FT_TH; }

}  // namespace orbx
