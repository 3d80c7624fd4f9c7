// orb_select.cu -- K3: per-level keypoint retention ("retainBest keeps ties") and the Harris response.
//
// Stages (v)-(vii) of OrbFeatureDetector::detect (reference src/FeatureExtractor.cpp:17; OpenCV
// KeyPointsFilter::retainBest and orb.cpp HarrisResponses):
//   retainBest(n): if count <= n keep all, else keep every keypoint whose response >= the n-th largest response.
//   HARRIS_SCORE (reference default): retainBest(2 q[l]) on the FAST score, Harris response, retainBest(q[l]) on it.
//   FAST_SCORE: retainBest(q[l]) on the FAST score.
// There is no spatial grid in cv::ORB; the quota is per pyramid level.
//
// k_select       : the FAST-score cut.  The n-th largest integer score is read off the 256-bin histogram that K2
//                  accumulated; every CTA recomputes it (256 adds) and then filters its stride of the candidate list.
// k_harris_select: one CTA per (level, frame).  Harris for each survivor (7x7 block of Sobel-like integer sums, then
//                  the 5 float operations in OpenCV's order, unfused), float k-th largest by rank counting
//                  (#strictly greater < q  <=>  response >= q-th largest), and the canonical (y, x) order of the
//                  kept points by rank counting on the packed coordinate.  All in shared memory.
#include "common.cuh"

namespace orbx {
namespace {

constexpr int SEL_THREADS = 256;
constexpr int SEL_CHUNKS = 8;
constexpr int HS_THREADS = 512;

__global__ void __launch_bounds__(SEL_THREADS)
k_select(const __grid_constant__ FrameGeom g, const Cand* __restrict__ cand, size_t cand_stride, Cand* __restrict__ surv,
         size_t surv_stride, FrameCounters* __restrict__ ctr, int32_t* __restrict__ fast_hint)
{
    __shared__ int s_thr;
    const int tid = threadIdx.x, lane = tid & 31;
    const int level = blockIdx.y, frame = blockIdx.z;
    const LevelGeom& L = g.lv[level];
    FrameCounters& C = ctr[frame];
    // what K2 found on this level, for K2's next launch: with more than 2.5 % of the interior pixels surviving NMS (the
    // Appendix-B stress frames: 3 %; camera-like frames: 0.1 %) its compass pre-test costs more than it saves (orb_fast.cu).
    // A speed hint only: the scores do not depend on it.
    if (fast_hint && frame == 0 && blockIdx.x == 0 && tid == 0)
        fast_hint[level] = (long long)C.ncand[level] * 40 > (long long)max(L.w - 62, 0) * max(L.h - 62, 0) ? 1 : 0;
    const int n = min(C.ncand[level], L.cand_cap);
    if (n == 0) return;
    const int target = g.score_type == ORBX_HARRIS_SCORE ? 2 * L.quota : L.quota;

    __shared__ uint32_t s_wsum[SEL_THREADS / 32];
    const uint32_t cnt = C.hist[level][tid];
    if (tid == 0) s_thr = (n <= target) ? 0 : 256;   // 0: keep everything; 256: keep nothing (target == 0)
    // S(s) = #candidates with score >= s (suffix sum of the histogram, non-increasing in s); the cut is the largest s with
    // S(s) >= target.  Warp-shuffle suffix scan + the totals of the warps above: 256 bins in a few steps.
    uint32_t S = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_down_sync(0xFFFFFFFFu, S, o);
        if (lane + o < 32) S += v;
    }
    if (lane == 0) s_wsum[tid >> 5] = S;
    __syncthreads();
    for (int wv = (tid >> 5) + 1; wv < SEL_THREADS / 32; wv++) S += s_wsum[wv];
    if (n > target && target > 0 && S >= (uint32_t)target && S - cnt < (uint32_t)target) s_thr = tid;
    __syncthreads();
    const uint32_t thr = (uint32_t)s_thr;

    const Cand* in = cand + frame * cand_stride + L.cand_off;
    Cand* out = surv + frame * surv_stride + L.surv_off;
    const int stride = SEL_CHUNKS * SEL_THREADS;
    for (int i0 = blockIdx.x * SEL_THREADS; i0 < n; i0 += stride) {
        const int i = i0 + tid;
        Cand c = { 0, 0 };
        bool keep = false;
        if (i < n) { c = in[i]; keep = c.score >= thr; }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&C.nsurv[level], __popc(bal));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (keep) {
                int pos = base + __popc(bal & ((1u << lane) - 1u));
                if (pos < L.surv_cap) out[pos] = c;
                else atomicOr(&C.overflow, 2);
            }
        }
    }
}

// OpenCV HarrisResponses, blockSize 7: a = sum Ix^2, b = sum Iy^2, c = sum Ix Iy over the 7x7 block around (x, y).
__device__ __forceinline__ float harris_response(const uint8_t* __restrict__ img, int pitch, int x, int y, float s4)
{
    int a = 0, b = 0, c = 0;
    const uint8_t* p = img + (size_t)(y - 4) * pitch + (x - 4);
    int r0[9], r1[9], r2[9];
#pragma unroll
    for (int i = 0; i < 9; i++) { r0[i] = __ldg(p + i); r1[i] = __ldg(p + pitch + i); }
#pragma unroll
    for (int j = 0; j < 7; j++) {
#pragma unroll
        for (int i = 0; i < 9; i++) r2[i] = __ldg(p + (size_t)(j + 2) * pitch + i);
#pragma unroll
        for (int i = 1; i <= 7; i++) {
            int Ix = (r1[i + 1] - r1[i - 1]) * 2 + (r0[i + 1] - r0[i - 1]) + (r2[i + 1] - r2[i - 1]);
            int Iy = (r2[i] - r0[i]) * 2 + (r2[i - 1] - r0[i - 1]) + (r2[i + 1] - r0[i + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
#pragma unroll
        for (int i = 0; i < 9; i++) { r0[i] = r1[i]; r1[i] = r2[i]; }
    }
    const float fa = (float)a, fb = (float)b, fc = (float)c;
    const float tr = __fadd_rn(fa, fb);
    const float det = __fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc));
    return __fmul_rn(__fsub_rn(det, __fmul_rn(__fmul_rn(0.04f, tr), tr)), s4);
}

// float -> unsigned key with the same order (zero canonicalised to +0)
__device__ __forceinline__ uint32_t ordered_key(float f)
{
    uint32_t u = __float_as_uint(f == 0.f ? 0.f : f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__global__ void __launch_bounds__(HS_THREADS)
k_harris_select(const __grid_constant__ FrameGeom g, const uint8_t* __restrict__ slots, size_t slot_stride,
                const Cand* __restrict__ surv, size_t surv_stride, Sel* __restrict__ sel, size_t sel_stride,
                FrameCounters* __restrict__ ctr, int smem_cap, float s4)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* s_key = reinterpret_cast<uint32_t*>(smem);                              // [smem_cap] ordered response keys
    uint32_t* s_xy = reinterpret_cast<uint32_t*>(smem + 4 * (size_t)smem_cap);        // [smem_cap]
    uint16_t* s_kept = reinterpret_cast<uint16_t*>(smem + 8 * (size_t)smem_cap);      // [smem_cap] indices of the kept points
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_prefix, s_remaining;
    __shared__ int s_nkept;

    const int tid = threadIdx.x, lane = tid & 31;
    const int level = blockIdx.x, frame = blockIdx.y;
    const LevelGeom& L = g.lv[level];
    FrameCounters& C = ctr[frame];
    const int m = min(min(C.nsurv[level], L.surv_cap), smem_cap);
    if (m == 0) { if (tid == 0) C.nsel[level] = 0; return; }

    const uint8_t* img = slots + frame * slot_stride + L.img_off;
    const Cand* in = surv + frame * surv_stride + L.surv_off;
    const bool harris = g.score_type == ORBX_HARRIS_SCORE;
    for (int i = tid; i < m; i += HS_THREADS) {
        const Cand c = in[i];
        s_xy[i] = c.xy;
        s_key[i] = ordered_key(harris ? harris_response(img, L.pitch, (int)(c.xy & 0xFFFFu), (int)(c.xy >> 16), s4) : (float)c.score);
    }
    if (tid == 0) { s_prefix = 0; s_remaining = (uint32_t)L.quota; s_nkept = 0; }
    __syncthreads();

    // ---- retainBest(q) on the response: T = key of the q-th largest, found by an 8-bit-per-pass radix select from the
    // top byte down; every key >= T is kept, so ties at the cut survive like in OpenCV.  (FAST_SCORE: the integer cut of
    // k_select already was the final one, everything is kept.)
    uint32_t T = 0;
    if (harris && L.quota == 0) T = 0xFFFFFFFFu;     // retainBest(0): nothing survives (no finite response maps to this key)
    else if (harris && m > L.quota) {
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += HS_THREADS) s_hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix, himask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
            for (int i = tid; i < m; i += HS_THREADS) {
                const uint32_t k = s_key[i];
                if ((k & himask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid < 32) {     // warp 0: the bin in which the count from the top reaches `remaining`
                const uint32_t remaining = s_remaining;
                uint32_t above = 0;         // keys of this pass's candidates in bins above the current group of 32
                for (int base = 224; base >= 0; base -= 32) {
                    const uint32_t cnt = s_hist[base + lane];
                    uint32_t suf = cnt;     // inclusive suffix sum over lanes lane .. 31
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t v = __shfl_down_sync(0xFFFFFFFFu, suf, o);
                        if (lane + o < 32) suf += v;
                    }
                    const bool hit = above + suf >= remaining && above + suf - cnt < remaining;
                    const uint32_t hitmask = __ballot_sync(0xFFFFFFFFu, hit);
                    if (hitmask) {
                        if (hit) {
                            s_prefix = prefix | ((uint32_t)(base + lane) << shift);
                            s_remaining = remaining - (above + suf - cnt);
                        }
                        break;
                    }
                    above += __shfl_sync(0xFFFFFFFFu, suf, 0);
                }
            }
            __syncthreads();
        }
        T = s_prefix;
    }

    // ---- compact the kept points, then order them by (y, x) with a rank count among the kept
    for (int i0 = 0; i0 < m; i0 += HS_THREADS) {
        const int i = i0 + tid;
        const bool keep = i < m && s_key[i] >= T;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&s_nkept, __popc(bal));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (keep) s_kept[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
    }
    __syncthreads();
    const int nk = s_nkept;
    Sel* out = sel + frame * sel_stride + L.sel_off;
    for (int a = tid; a < nk; a += HS_THREADS) {
        const int i = s_kept[a];
        const uint32_t key = s_xy[i];
        int rank = 0;
        for (int b = 0; b < nk; b++) rank += (s_xy[s_kept[b]] < key);
        Sel s;
        s.xy = key;
        s.response = key_value(s_key[i]);
        out[rank] = s;      // rank < nk <= m <= surv_cap == capacity of the selected list
    }
    if (tid == 0) C.nsel[level] = nk;
}

}  // namespace

cudaError_t launch_select(const FrameGeom& g, const Cand* cand, size_t cand_stride, Cand* surv, size_t surv_stride,
                          FrameCounters* ctr, int nframes, int32_t* fast_hint, cudaStream_t s)
{
    dim3 grid(SEL_CHUNKS, g.nlevels, nframes);
    k_select<<<grid, SEL_THREADS, 0, s>>>(g, cand, cand_stride, surv, surv_stride, ctr, fast_hint);
    return cudaGetLastError();
}

size_t harris_select_smem(int max_surv_cap) { return (size_t)max_surv_cap * 10 + 16; }

cudaError_t harris_select_prepare(int max_surv_cap)
{
    return cudaFuncSetAttribute(k_harris_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)harris_select_smem(max_surv_cap));
}

cudaError_t launch_harris_select(const FrameGeom& g, const uint8_t* slots, size_t slot_stride, const Cand* surv,
                                 size_t surv_stride, Sel* sel, size_t sel_stride, FrameCounters* ctr, int nframes,
                                 int max_surv_cap, float s4, cudaStream_t s)
{
    dim3 grid(g.nlevels, nframes);
    k_harris_select<<<grid, HS_THREADS, harris_select_smem(max_surv_cap), s>>>(g, slots, slot_stride, surv, surv_stride, sel,
                                                                                sel_stride, ctr, max_surv_cap, s4);
    return cudaGetLastError();
}

}  // namespace orbx
