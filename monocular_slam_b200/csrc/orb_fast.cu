// orb_fast.cu -- K2: FAST-9/16 detection, integer corner score, 3x3 strict non-max suppression, border filter,
// compaction -- every pyramid level of every frame of the batch in ONE launch.
//
// Stages (iii)-(iv) of OrbFeatureDetector::detect (reference src/FeatureExtractor.cpp:17; OpenCV fast.cpp /
// fast_score.cpp):
//   d[k] = I(p) - I(ring k);  m = max over the 16 arcs of 9 contiguous ring pixels of max(min d, min -d)
//   corner  <=>  m > threshold (20);  score = m - 1;  keep iff score > all 8 neighbours' scores (non-corners score 0)
//   and 31 <= x < w-31, 31 <= y < h-31.
//
// A CTA owns a 124x30 tile of the interior [31,w-31) x [31,h-31) of one level.  It stages the tile plus a 4-pixel
// halo in shared memory with 16-byte loads (rows are 128-byte pitched), then
//   1. scores the (tile + 1) region branch-free, four horizontally adjacent pixels per thread and two pixels per
//      instruction: ring-minus-centre differences live in packed s16x2 lanes (VIADD.16x2) and the sliding
//      9-of-16 window minimum / maximum is a log-step network of VIMNMX.S16x2 (55 min/max per sign and pixel pair),
//      which yields m = max(max_k min9(d), -min_k max9(d)) exactly -- no corner pre-test, no divergence.  A warp
//      covers one score row (32 groups of 4 pixels), 21 aligned 32-bit shared-memory loads feed 4 pixels;
//   2. NMS on the score tile (a whole word of four zero scores is skipped at once); survivors are staged in shared
//      memory, the CTA reserves its range of the level's candidate list with ONE global atomic and writes
//      (x, y, score); the per-level score histogram used by the retainBest cut (K3) gets one RED per survivor.
// Bound: integer ALU issue (VIMNMX/PRMT), not HBM: each level byte is read once from HBM and ~1.4x from L2.
#include "common.cuh"

namespace orbx {
namespace {

constexpr int FT_TW = 124;
constexpr int FT_TH = 30;
constexpr int FT_THREADS = 256;
constexpr int FT_SP = 160;               // smem pitch (image and score share column coordinates): <= 15 + 4 + 124 + 4
constexpr int FT_SH = FT_TH + 8;         // 38 image rows
constexpr int FT_CH = FT_TH + 2;         // 32 score rows = 8 warps x 4
constexpr int FT_EMIT = (FT_TW / 2 + 1) * (FT_TH / 2 + 1);   // NMS allows at most one maximum per 2x2 block

// 16x2 lanes from bytes (0,1) / (2,3) of g
__device__ __forceinline__ uint32_t lanes_lo(uint32_t g) { return __byte_perm(g, 0u, 0x4140); }
__device__ __forceinline__ uint32_t lanes_hi(uint32_t g) { return __byte_perm(g, 0u, 0x4342); }

// bytes [4 + dx, 8 + dx) of the 12-byte window {w0, w1, w2} (w1 holds the four centre columns)
template <int DX>
__device__ __forceinline__ uint32_t window4(uint32_t w0, uint32_t w1, uint32_t w2)
{
    if (DX == 0) return w1;
    if (DX == -1) return __byte_perm(w0, w1, 0x6543);
    if (DX == -2) return __byte_perm(w0, w1, 0x5432);
    if (DX == -3) return __byte_perm(w0, w1, 0x4321);
    if (DX == 1) return __byte_perm(w1, w2, 0x4321);
    if (DX == 2) return __byte_perm(w1, w2, 0x5432);
    return __byte_perm(w1, w2, 0x6543);   // DX == 3
}

// m + 256, m = max over the 16 arcs of 9 contiguous ring positions of max(min d, min -d), for two pixels at once.
// d[] holds the BIASED differences I(ring) - I(centre) + 256 in s16x2 lanes: every lane stays in [1, 511], so the
// per-lane additions that produce them are plain 32-bit adds (no carry between lanes, and ptxas may place them on the
// FMA pipe as IMAD.IADD), while the min/max network is shift-invariant.
__device__ __forceinline__ uint32_t arc_strength2(const uint32_t (&d)[16])
{
    uint32_t lo2[8], hi2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {       // window {2j+1, 2j+2}
        lo2[j] = __vmins2(d[2 * j + 1], d[(2 * j + 2) & 15]);
        hi2[j] = __vmaxs2(d[2 * j + 1], d[(2 * j + 2) & 15]);
    }
    uint32_t lo4[8], hi4[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {       // window 2j+1 .. 2j+4
        lo4[j] = __vmins2(lo2[j], lo2[(j + 1) & 7]);
        hi4[j] = __vmaxs2(hi2[j], hi2[(j + 1) & 7]);
    }
    uint32_t best_lo = 0u, best_hi = 0x7FFF7FFFu;   // running max of arc minima / min of arc maxima
#pragma unroll
    for (int j = 0; j < 8; j++) {       // window 2j+1 .. 2j+8, extended by 2j (arc 2j) or 2j+9 (arc 2j+1)
        const uint32_t lo8 = __vmins2(lo4[j], lo4[(j + 2) & 7]);
        const uint32_t hi8 = __vmaxs2(hi4[j], hi4[(j + 2) & 7]);
        best_lo = __vmaxs2(best_lo, __vmins2(lo8, d[2 * j]));
        best_lo = __vmaxs2(best_lo, __vmins2(lo8, d[(2 * j + 9) & 15]));
        best_hi = __vmins2(best_hi, __vmaxs2(hi8, d[2 * j]));
        best_hi = __vmins2(best_hi, __vmaxs2(hi8, d[(2 * j + 9) & 15]));
    }
    // bright: best_lo - 256; dark: 256 - best_hi; both re-biased by +256 (lanes of 512 - best_hi are in [1, 511])
    return __vmaxs2(best_lo, 0x02000200u - best_hi);
}

// biased strength M = m + 256 -> score m - 1 if m > thr else 0
__device__ __forceinline__ uint32_t score_from_strength(uint32_t M, int thr) { return (int)M > thr + 256 ? M - 257u : 0u; }

__global__ void __launch_bounds__(FT_THREADS)
k_fast(const __grid_constant__ FrameGeom g, const uint8_t* __restrict__ slots, size_t slot_stride, Cand* __restrict__ cand,
       size_t cand_stride, FrameCounters* __restrict__ ctr)
{
    __shared__ __align__(16) uint8_t s_img[FT_SH * FT_SP];
    __shared__ __align__(16) uint8_t s_score[FT_CH * FT_SP];
    __shared__ Cand s_emit[FT_EMIT];
    __shared__ int s_en, s_base;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int frame = blockIdx.y;

    int level = 0;
    while (level + 1 < g.nlevels && (int)blockIdx.x >= g.lv[level + 1].tile_start) level++;
    const LevelGeom& L = g.lv[level];
    const int t = blockIdx.x - L.tile_start;
    const int ty = t / L.tiles_x, tx = t - ty * L.tiles_x;
    const int w = L.w, h = L.h, pitch = L.pitch;
    const uint8_t* img = slots + frame * slot_stride + L.img_off;
    const int ox = 31 + tx * FT_TW, oy = 31 + ty * FT_TH;   // first output pixel of the tile
    const int gx0 = (ox - 4) & ~15, gy0 = oy - 4;           // image coordinates of s_img[0][0]
    const int thr = g.fast_threshold;

    if (tid == 0) s_en = 0;
    for (int i = tid; i < FT_SH * (FT_SP / 16); i += FT_THREADS) {
        int r = i / (FT_SP / 16), v = i - r * (FT_SP / 16);
        int gy = gy0 + r, gx = gx0 + v * 16;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (gy < h && gx < pitch) val = *reinterpret_cast<const uint4*>(img + (size_t)gy * pitch + gx);
        *reinterpret_cast<uint4*>(s_img + r * FT_SP + v * 16) = val;
    }
    __syncthreads();

    // ---- 1: scores of rows oy-1 .. oy+30, columns ox-3 .. ox+124 (32 groups of 4, group start is a multiple of 4)
    const int c = (ox - 3 - gx0) + 4 * lane;   // smem column of the group's first pixel (multiple of 4)
#pragma unroll 1
    for (int it = 0; it < FT_CH / 8; it++) {
        const int sr = wid + 8 * it;            // score row; its centre pixels sit in image smem row sr + 3
        const uint32_t* base = reinterpret_cast<const uint32_t*>(s_img + (sr + 3) * FT_SP + c - 4);
        auto W = [&](int dy, int i) { return base[dy * (FT_SP / 4) + i]; };
        // biased ring differences d[k] = I(ring k) - I(centre) + 256 for pixel pairs A = (c, c+1) and B = (c+2, c+3)
        const uint32_t ctr4 = W(0, 1);
        const uint32_t cA = 0x01000100u - lanes_lo(ctr4), cB = 0x01000100u - lanes_hi(ctr4);   // 256 - centre per lane
        uint32_t dA[16], dB[16];
#define RING(K, DX, DY)                                                    \
        {                                                                  \
            const uint32_t g4 = window4<DX>(W(DY, 0), W(DY, 1), W(DY, 2)); \
            dA[K] = lanes_lo(g4) + cA;                                     \
            dB[K] = lanes_hi(g4) + cB;                                     \
        }
        RING(0, 0, 3)   RING(1, 1, 3)    RING(2, 2, 2)    RING(3, 3, 1)
        RING(4, 3, 0)   RING(5, 3, -1)   RING(6, 2, -2)   RING(7, 1, -3)
        RING(8, 0, -3)  RING(9, -1, -3)  RING(10, -2, -2) RING(11, -3, -1)
        RING(12, -3, 0) RING(13, -3, 1)  RING(14, -2, 2)  RING(15, -1, 3)
#undef RING
        const uint32_t mA = arc_strength2(dA), mB = arc_strength2(dB);
        const uint32_t s0 = score_from_strength(mA & 0xFFFFu, thr);
        const uint32_t s1 = score_from_strength(mA >> 16, thr);
        const uint32_t s2 = score_from_strength(mB & 0xFFFFu, thr);
        const uint32_t s3 = score_from_strength(mB >> 16, thr);
        *reinterpret_cast<uint32_t*>(s_score + sr * FT_SP + c) = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
    }
    __syncthreads();

    // ---- 2: non-max suppression on the tile proper (score row yy + 1, smem column ox - gx0 + xx)
    const int cx0 = ox - gx0;                         // multiple of 4 plus 3: tile columns start inside a word
    const int word0 = (cx0 & ~3), nwords = ((cx0 + FT_TW - 1) >> 2) - (cx0 >> 2) + 1;
    for (int i = tid; i < nwords * FT_TH; i += FT_THREADS) {
        const int yy = i / nwords, wi = i - yy * nwords;
        const int col = word0 + 4 * wi;
        const uint32_t word = *reinterpret_cast<const uint32_t*>(s_score + (yy + 1) * FT_SP + col);
        if (word == 0) continue;
        const int y = oy + yy;
        if (y >= h - 31) continue;
#pragma unroll
        for (int bq = 0; bq < 4; bq++) {
            const int s = (word >> (8 * bq)) & 0xFF;
            const int cc = col + bq;
            const int x = gx0 + cc;
            if (s == 0 || cc < cx0 || cc >= cx0 + FT_TW || x >= w - 31) continue;
            const uint8_t* sp = s_score + (yy + 1) * FT_SP + cc;
            if (s > sp[-1] && s > sp[1] && s > sp[-FT_SP - 1] && s > sp[-FT_SP] && s > sp[-FT_SP + 1] && s > sp[FT_SP - 1] &&
                s > sp[FT_SP] && s > sp[FT_SP + 1]) {
                int pos = atomicAdd(&s_en, 1);
                Cand cnd;
                cnd.xy = ((uint32_t)y << 16) | (uint32_t)x;
                cnd.score = (uint32_t)s;
                s_emit[pos] = cnd;
            }
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
