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
// halo in shared memory with 8-byte loads (rows are 128-byte pitched), widening every pixel to 16 bits on the way so
// that two horizontally adjacent pixels already form one packed s16x2 register operand, then
//   1. scores the (tile + 1) region branch-free, four horizontally adjacent pixels per thread and two pixels per
//      instruction: the RAW ring pixels live in packed 16x2 lanes and the sliding 9-of-16 window minimum / maximum
//      is a network of 36 min/max per sign and pixel pair -- the 16 (min, max) pairs with shared inputs as HFMA2.RELU /
//      HFMA2 on the FMA pipe, the rest as VIMNMX(3).S16x2 on the ALU pipe; min/max is shift-invariant, so the
//      centre is subtracted once from the two results instead of from the 16 ring pixels, which yields
//      m = max(max_k min9(ring) - c, c - min_k max9(ring)) exactly -- no corner pre-test, no divergence.  A warp
//      covers one score row (32 groups of 4 pixels); 21 aligned 64-bit shared-memory loads and 18 byte-permutes feed
//      4 pixels (ring columns at even offsets are register operands as loaded);
//   2. NMS on the score tile, again four pixels per thread in packed u16x2 lanes: each warp walks its rows with a rolling
//      3-row window of horizontal maxima, so the 8-neighbour maximum costs one 3-input VIMNMX per pixel pair, and
//      "strictly greater" is one subtraction + bit test; survivors are compacted per row with two warp ballots (a lane
//      keeps at most 2 of its 4 pixels) into a shared-memory list that reuses the image tile, the CTA reserves its range
//      of the level's candidate list with ONE global atomic and writes (x, y, score); the per-level score histogram used
//      by the retainBest cut (K3) gets one RED per survivor.
// CTAs are 128 threads (4 warps x 8 rows): small barrier domains, 8 CTAs per SM.
// Bound: instruction issue (ncu: issue slots 80 % busy, ALU pipe 71 %), not HBM: each level byte is read once from HBM
// and ~1.4x from L2.  About 95 instructions per pixel in total.
#include "common.cuh"

namespace orbx {
namespace {

constexpr int FT_TW = 124;
constexpr int FT_TH = 30;
#ifndef FT_THREADS_N
#define FT_THREADS_N 128
#endif
#ifndef FT_MINB
#define FT_MINB 8
#endif
constexpr int FT_THREADS = FT_THREADS_N;
constexpr int FT_NW = FT_THREADS / 32;
constexpr int FT_SP = 160;               // smem pitch (image and score share column coordinates): <= 15 + 4 + 124 + 4
constexpr int FT_SH = FT_TH + 8;         // 38 image rows
constexpr int FT_CH = FT_TH + 2;         // 32 score rows = 4 warps x 8
constexpr int FT_EMIT = (FT_TW / 2 + 1) * (FT_TH / 2 + 1);   // NMS allows at most one maximum per 2x2 block

// 16x2 lanes from bytes (0,1) / (2,3) of g
__device__ __forceinline__ uint32_t lanes_lo(uint32_t g) { return __byte_perm(g, 0u, 0x4140); }
__device__ __forceinline__ uint32_t lanes_hi(uint32_t g) { return __byte_perm(g, 0u, 0x4342); }
// lanes (hi of a, lo of b): the pixel pair that straddles two aligned pairs
__device__ __forceinline__ uint32_t straddle(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x5432); }

// The 16-bit lanes hold integers 0..255; read as fp16 they are subnormals whose order and sums are those of the
// integers, so (min, max) of two lane registers can be had on the otherwise idle FMA pipe, exactly, as
// t = relu(a - b), min = a - t, max = b + t (3 HFMA2 for both results; tools/fastnet_probe.cu checks it exhaustively).
__device__ __forceinline__ uint32_t hfma2_relu(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ void minmax_fma(uint32_t a, uint32_t b, uint32_t& mn, uint32_t& mx)
{
    const uint32_t one = 0x3C003C00u, neg_one = 0xBC00BC00u;
    const uint32_t t = hfma2_relu(b, neg_one, a);
    mn = hfma2(t, neg_one, a);
    mx = hfma2(t, one, b);
}

// thr-clamped corner strength of two pixels at once.  r[] holds the RAW ring pixels (0..255) in 16x2 lanes, c the centre
// pixels.  Sliding min/max is shift-invariant, so the centre is subtracted once at the end instead of from every ring
// pixel:  m = max(max_arcs min_arc(ring) - c, c - min_arcs max_arc(ring)).
// Arc k = ring positions k .. k+8.  With W_j = positions 2j+1 .. 2j+8 (min of two 4-windows), arcs 2j and 2j+1 are
// W_j extended by position 2j resp. 2j+9, and max(min(W,a), min(W,b)) == min(W, max(a,b)), so both arcs cost one
// 2-input and one 3-input instruction: 36 min/max per sign and pixel pair (8 + 8 + 8 + 8 + 4).  The 16 (min, max)
// pairs that share their inputs (first level, arc extensions) run on the FMA pipe (48 HFMA2), the other 40 + 4
// instructions are VIMNMX(3).S16x2 on the ALU pipe, which is the pipe this kernel is bound by.
// Returns max(m - thr, 0) per lane; K = (thr + 256) in both lanes.
__device__ __forceinline__ uint32_t arc_strength2(const uint32_t (&r)[16], uint32_t c, uint32_t K)
{
    uint32_t lo2[8], hi2[8];
#pragma unroll
    for (int j = 0; j < 8; j++) minmax_fma(r[2 * j + 1], r[(2 * j + 2) & 15], lo2[j], hi2[j]);   // window {2j+1, 2j+2}
    uint32_t lo4[8], hi4[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {       // window 2j+1 .. 2j+4
        lo4[j] = __vmins2(lo2[j], lo2[(j + 1) & 7]);
        hi4[j] = __vmaxs2(hi2[j], hi2[(j + 1) & 7]);
    }
    uint32_t alo[8], ahi[8];            // best of arcs 2j and 2j+1
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t emn, emx;
        minmax_fma(r[2 * j], r[(2 * j + 9) & 15], emn, emx);
        alo[j] = __vimin3_s16x2(lo4[j], lo4[(j + 2) & 7], emx);
        ahi[j] = __vimax3_s16x2(hi4[j], hi4[(j + 2) & 7], emn);
    }
    const uint32_t best_lo = __vimax3_s16x2(__vimax3_s16x2(alo[0], alo[1], alo[2]), __vimax3_s16x2(alo[3], alo[4], alo[5]),
                                            __vmaxs2(alo[6], alo[7]));
    const uint32_t best_hi = __vimin3_s16x2(__vimin3_s16x2(ahi[0], ahi[1], ahi[2]), __vimin3_s16x2(ahi[3], ahi[4], ahi[5]),
                                            __vmins2(ahi[6], ahi[7]));
    // bright: best_lo - c, dark: c - best_hi, both biased by +256 so every lane stays in [1, 511] (plain 32-bit adds)
    const uint32_t bright = best_lo + 0x01000100u - c, dark = c + 0x01000100u - best_hi;
    return __vimax3_s16x2(bright, dark, K) - K;
}

__global__ void __launch_bounds__(FT_THREADS, FT_MINB)
k_fast(const __grid_constant__ FrameGeom g, const uint8_t* __restrict__ slots, size_t slot_stride, Cand* __restrict__ cand,
       size_t cand_stride, FrameCounters* __restrict__ ctr)
{
    __shared__ __align__(16) uint16_t s_img[FT_SH * FT_SP];       // pixels widened to 16 bits
    __shared__ __align__(16) uint16_t s_score[FT_CH * FT_SP];     // m - thr for corners (1 .. 255 - thr), 0 otherwise
    // the survivor list reuses the image tile: no thread reads s_img after the barrier that ends the scoring pass
    static_assert(sizeof(Cand) * FT_EMIT <= sizeof(uint16_t) * FT_SH * FT_SP, "survivor list must fit in the image tile");
    Cand* s_emit = reinterpret_cast<Cand*>(s_img);
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
    // s_score needs no clearing: every entry NMS consumes for a pixel of the tile proper is written by the scoring pass
    // 8-byte units: 38 rows x 20 units = 760 = 3 per thread (less 8), so every warp reaches the barrier with the same work
#pragma unroll
    for (int k = 0; k < (FT_SH * (FT_SP / 8) + FT_THREADS - 1) / FT_THREADS; k++) {
        const int i = tid + k * FT_THREADS;
        if (i < FT_SH * (FT_SP / 8)) {
            const int r = i / (FT_SP / 8), v = i - r * (FT_SP / 8);
            const int gy = gy0 + r, gx = gx0 + v * 8;
            uint2 val = make_uint2(0, 0);
            if (gy < h && gx < pitch) val = *reinterpret_cast<const uint2*>(img + (size_t)gy * pitch + gx);
            *reinterpret_cast<uint4*>(s_img + r * FT_SP + v * 8) = make_uint4(lanes_lo(val.x), lanes_hi(val.x), lanes_lo(val.y), lanes_hi(val.y));
        }
    }
    __syncthreads();

    // ---- 1: scores of rows oy-1 .. oy+30, columns ox-3 .. ox+124 (32 groups of 4, group start is a multiple of 4)
    const int c = (ox - 3 - gx0) + 4 * lane;   // smem column of the group's first pixel (multiple of 4)
    const uint32_t K = (uint32_t)(thr + 256) * 0x00010001u;
#pragma unroll 1
    for (int it = 0; it < FT_CH / FT_NW; it++) {
        const int sr = wid + FT_NW * it;            // score row; its centre pixels sit in image smem row sr + 3
        if (oy - 1 + sr > h - 31) continue;         // below the last row NMS can look at (bottom tiles of a level); warp-uniform
        // pixel pairs of row dy: q[i] = pixels (c - 4 + 2i, c - 3 + 2i), i = 0..5, as s16x2 registers
        const uint2* base = reinterpret_cast<const uint2*>(s_img + (sr + 3) * FT_SP + c - 4);
        auto P = [&](int dy, int i) { return base[dy * (FT_SP / 4) + i]; };
        uint32_t rA[16], rB[16];   // raw ring pixels for pixel pairs A = (c, c+1), B = (c+2, c+3)
        uint32_t cA, cB;           // centre pixels
        {
            const uint2 p0 = P(0, 0), p1 = P(0, 1), p2 = P(0, 2);          // row y: centre, ring 4 (dx +3), ring 12 (dx -3)
            cA = p1.x; cB = p1.y;
            rA[4] = straddle(p1.y, p2.x);  rB[4] = straddle(p2.x, p2.y);
            rA[12] = straddle(p0.x, p0.y); rB[12] = straddle(p0.y, p1.x);
#define ROW3(DY, KM, K0, KP)   /* rows y +- 3: dx -1, 0, +1 */                                   \
            {                                                                                     \
                const uint2 a0 = P(DY, 0), a1 = P(DY, 1), a2 = P(DY, 2);                          \
                const uint32_t m = straddle(a0.y, a1.x), z = straddle(a1.x, a1.y), n = straddle(a1.y, a2.x); \
                rA[KM] = m;    rB[KM] = z;                                                        \
                rA[K0] = a1.x; rB[K0] = a1.y;                                                     \
                rA[KP] = z;    rB[KP] = n;                                                        \
            }
#define ROW2(DY, KM, KP)       /* rows y +- 2: dx -2, +2 */                                       \
            {                                                                                     \
                const uint2 a0 = P(DY, 0), a1 = P(DY, 1), a2 = P(DY, 2);                          \
                rA[KM] = a0.y; rB[KM] = a1.x;                                                     \
                rA[KP] = a1.y; rB[KP] = a2.x;                                                     \
            }
#define ROW1(DY, KM, KP)       /* rows y +- 1: dx -3, +3 */                                       \
            {                                                                                     \
                const uint2 a0 = P(DY, 0), a1 = P(DY, 1), a2 = P(DY, 2);                          \
                rA[KM] = straddle(a0.x, a0.y); rB[KM] = straddle(a0.y, a1.x);                     \
                rA[KP] = straddle(a1.y, a2.x); rB[KP] = straddle(a2.x, a2.y);                     \
            }
            ROW3(3, 15, 0, 1)    // ring 15 (-1,+3), 0 (0,+3), 1 (+1,+3)
            ROW2(2, 14, 2)       // ring 14 (-2,+2), 2 (+2,+2)
            ROW1(1, 13, 3)       // ring 13 (-3,+1), 3 (+3,+1)
            ROW1(-1, 11, 5)      // ring 11 (-3,-1), 5 (+3,-1)
            ROW2(-2, 10, 6)      // ring 10 (-2,-2), 6 (+2,-2)
            ROW3(-3, 9, 8, 7)    // ring 9 (-1,-3), 8 (0,-3), 7 (+1,-3)
#undef ROW3
#undef ROW2
#undef ROW1
        }
        // store m - thr for corners, 0 otherwise (order-preserving, so NMS is unaffected)
        const uint32_t sA = arc_strength2(rA, cA, K), sB = arc_strength2(rB, cB, K);
        *reinterpret_cast<uint2*>(s_score + sr * FT_SP + c) = make_uint2(sA, sB);
    }
    __syncthreads();

    // ---- 2: non-max suppression on the tile proper.  Warp w owns the output rows [RPW*w, RPW*(w+1)) of score rows 1 .. 30
    // and walks them top to bottom with a rolling 3-row window of horizontal maxima, four pixels per lane in packed
    // u16x2 lanes (same lane -> column mapping as above):  per row, hx = max(left, right) and h3 = max(hx, centre);
    // the 8-neighbour maximum of row y is max3(hx[y], h3[y-1], h3[y+1]).
    {
        constexpr int RPW = FT_CH / FT_NW;
        const int cx0 = ox - gx0;                     // smem column of the tile's first pixel
        uint32_t colmask = 0;                          // which of this lane's four columns belong to the tile and the interior
#pragma unroll
        for (int bq = 0; bq < 4; bq++) {
            const int cc = c + bq;
            if (cc >= cx0 && cc < cx0 + FT_TW && gx0 + cc < w - 31) colmask |= 1u << bq;
        }
        struct Row { uint32_t mA, mB, hxA, hxB, h3A, h3B; };
        auto load_row = [&](int sr) {
            const uint16_t* rowp = s_score + sr * FT_SP + c;
            const uint2 mid = *reinterpret_cast<const uint2*>(rowp);             // pairs A = (c, c+1), B = (c+2, c+3)
            const uint32_t l = *reinterpret_cast<const uint32_t*>(rowp - 2), r = *reinterpret_cast<const uint32_t*>(rowp + 4);
            const uint32_t z = straddle(mid.x, mid.y);
            Row R;
            R.mA = mid.x; R.mB = mid.y;
            R.hxA = __vmaxu2(straddle(l, mid.x), z);
            R.hxB = __vmaxu2(z, straddle(mid.y, r));
            R.h3A = __vmaxu2(R.hxA, mid.x);
            R.h3B = __vmaxu2(R.hxB, mid.y);
            return R;
        };
        const int r_begin = max(RPW * wid, 1), r_end = min(RPW * wid + RPW, FT_CH - 1);
        Row prev = load_row(r_begin - 1), cur = load_row(r_begin);
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll 2
        for (int sr = r_begin; sr < r_end; sr++) {
            const Row next = load_row(sr + 1);
            const int y = oy - 1 + sr;
            const uint32_t nA = __vimax3_u16x2(cur.hxA, prev.h3A, next.h3A), nB = __vimax3_u16x2(cur.hxB, prev.h3B, next.h3B);
            // strictly greater than all 8 neighbours  <=>  score - nmax >= 1  <=>  bit 8 of (score + 256 - nmax - 1);
            // every lane stays in [0, 510], so the 32-bit subtractions never borrow across lanes
            const uint32_t fA = (cur.mA | 0x01000100u) - nA - 0x00010001u, fB = (cur.mB | 0x01000100u) - nB - 0x00010001u;
            uint32_t km = ((fA >> 8) & 1u) | ((fA >> 23) & 2u) | ((fB >> 6) & 4u) | ((fB >> 21) & 8u);
            km = (y < h - 31) ? (km & colmask) : 0u;
            // two horizontally adjacent pixels cannot both be strict maxima: a lane keeps at most 2 of its 4 pixels
            const int n = __popc(km);
            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, n >= 1);
            if (b0) {                                  // uniform over the warp
                const uint32_t b1 = __ballot_sync(0xFFFFFFFFu, n >= 2);
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&s_en, __popc(b0) + __popc(b1));
                pos = __shfl_sync(0xFFFFFFFFu, pos, 0) + __popc(b0 & lt) + __popc(b1 & lt);
                uint32_t m = km;
                while (m) {
                    const int bq = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t pair = (bq & 2) ? cur.mB : cur.mA;
                    Cand cnd;
                    cnd.xy = ((uint32_t)y << 16) | (uint32_t)(gx0 + c + bq);
                    cnd.score = ((pair >> (16 * (bq & 1))) & 0xFFFFu) + (uint32_t)thr - 1u;   // FAST score = m - 1
                    s_emit[pos++] = cnd;
                }
            }
            prev = cur;
            cur = next;
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
