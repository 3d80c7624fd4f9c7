// hamming.cu -- K6: brute-force 256-bit Hamming kNN(k=2) + Lowe ratio test for sm_100a.
//
// Replaces BFMatcher(NORM_HAMMING,false).knnMatch(d1,d2,raw,2) and the ratio loop of
// matchFeatures (reference src/CameraPoseEstimator.cpp:200-213).  The distance is the one spelled out in
// ThirdParty/DBoW2/DBoW2/FORB.cpp:81-101: XOR of the two 256-bit strings, population count, summed.
//
// Layout.  Descriptors are rows of 32 bytes (2 x uint4).  A CTA of 128 threads owns 256 query rows, two per thread,
// held in registers for the whole kernel.  The train rows of the CTA's split are streamed through a 4-stage ring of
// 4 KB shared-memory tiles filled by the TMA engine (cp.async.bulk + mbarrier complete_tx); every lane reads the
// same train row (a shared-memory broadcast) and XORs it against its two queries.  A plain popcount needs 8 POPC per
// comparison and is bound by the POPC pipe (16 lanes/clk/SM); ham256() first compresses the XOR words with carry-save
// adders on the 4x wider LOP3 pipe, which leaves 5 POPC per comparison and balances the two pipes (DESIGN.md "K6").
//
// Tie rule.  OpenCV inserts candidates in ascending train order with a strict `<`, i.e. the result is the two
// lexicographically smallest (distance, trainIdx) pairs.  Packing key = distance << 23 | trainIdx makes that a
// plain unsigned min: best1 = min(best1, max(best0, key)); best0 = min(best0, key).  Train sets larger than 2^23
// rows are processed in chunks by the host and merged with the same rule on 64-bit keys.
#include <stdlib.h>
#include <algorithm>

#include "common.cuh"

namespace orbx {
namespace {

constexpr int HT_THREADS = 128;
constexpr int HT_QPT = 2;
constexpr int HT_QB = HT_THREADS * HT_QPT;   // queries per CTA
constexpr int HT_TT = 128;                   // train rows per stage
constexpr int HT_STAGES = 4;
constexpr int HT_IDX_BITS = 23;
constexpr uint32_t HT_IDX_MASK = (1u << HT_IDX_BITS) - 1;
constexpr uint32_t HT_NONE = 0xFFFFFFFFu;    // distance field 511: larger than any real key

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void top2_insert(uint32_t& b0, uint32_t& b1, uint32_t key)
{
    uint32_t hi = max(b0, key);
    b0 = min(b0, key);
    b1 = min(b1, hi);
}

// Carry-save adder on 32 bit positions at once: a + b + c == s + 2 * cy (two LOP3: 3-input XOR and majority).
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& s, uint32_t& cy)
{
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s) : "r"(a), "r"(b), "r"(c));
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(cy) : "r"(a), "r"(b), "r"(c));
}

#ifndef HT_CSA
#define HT_CSA 1    // measured on B200 (1 M x 125 k): 0 -> 549, 1 -> 833, 2 -> 812 Gcmp/s
#endif
// 256-bit Hamming distance (ThirdParty/DBoW2/DBoW2/FORB.cpp:81-101: XOR, population count, sum).  The POPC pipe issues
// 16 lanes/clk/SM against 64 for LOP3, so instead of 8 POPC the eight XOR words are first compressed with carry-save
// adders (Harley-Seal): x0..x6 -> s3 + 2 (c1 + c2 + c3) -> s3 + 2 s5 + 4 c5, leaving 4 POPC (HT_CSA == 2) or 5 (== 1).
__device__ __forceinline__ uint32_t ham256(const uint32_t (&q)[8], const uint4& a, const uint4& b)
{
    const uint32_t x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x2 = q[2] ^ a.z, x3 = q[3] ^ a.w;
    const uint32_t x4 = q[4] ^ b.x, x5 = q[5] ^ b.y, x6 = q[6] ^ b.z, x7 = q[7] ^ b.w;
#if HT_CSA == 0
    return __popc(x0) + __popc(x1) + __popc(x2) + __popc(x3) + __popc(x4) + __popc(x5) + __popc(x6) + __popc(x7);
#else
    uint32_t s1, c1, s2, c2, s3, c3;
    csa(x0, x1, x2, s1, c1);
    csa(x3, x4, x5, s2, c2);
    csa(s1, s2, x6, s3, c3);
#if HT_CSA == 1
    return __popc(s3) + __popc(x7) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
#else
    uint32_t s5, c5;
    csa(c1, c2, c3, s5, c5);
    return __popc(s3) + __popc(x7) + 2 * __popc(s5) + 4 * __popc(c5);
#endif
#endif
}

__device__ __forceinline__ hamx_top2 decode_top2(uint32_t k0, uint32_t k1, int64_t offset)
{
    hamx_top2 r;
    r.dist0 = k0 == HT_NONE ? -1 : (int32_t)(k0 >> HT_IDX_BITS);
    r.idx0 = k0 == HT_NONE ? -1 : (int32_t)((int64_t)(k0 & HT_IDX_MASK) + offset);
    r.dist1 = k1 == HT_NONE ? -1 : (int32_t)(k1 >> HT_IDX_BITS);
    r.idx1 = k1 == HT_NONE ? -1 : (int32_t)((int64_t)(k1 & HT_IDX_MASK) + offset);
    return r;
}

__device__ __forceinline__ unsigned long long top2_key64(int32_t d, int32_t i)
{
    return i < 0 ? ~0ull : ((unsigned long long)(uint32_t)d << 32) | (uint32_t)i;
}

__device__ __forceinline__ void top2_insert64(unsigned long long& b0, unsigned long long& b1, unsigned long long key)
{
    unsigned long long hi = b0 > key ? b0 : key;
    b0 = b0 < key ? b0 : key;
    b1 = b1 < hi ? b1 : hi;
}

// ---- train-sharded matching over peer memory (one process per GPU, NVLink P2P) ----------------------------------------
// Every rank owns a gather buffer [2 parities][world][nq_max] of hamx_top2 plus [2][world] epoch flags, mapped into all
// peers (cudaIpc).  The matching kernel's epilogue stores each query's local top-2 straight into slot [rank] of EVERY
// rank's gather buffer (16-byte stores over NVLink; no staging copy, no separate all-gather launch); the last CTA to
// finish publishes the call's epoch in every rank's flag slot with a system-scope release.  k_merge_top2_p2p waits for
// the world's flags with system-scope acquires and reduces the slots with the lexicographic rule.
constexpr int HT_MAX_WORLD = 16;
struct P2PView {
    hamx_top2* gather[HT_MAX_WORLD];     // base of rank r's gather buffer as mapped into this process
    unsigned int* flags[HT_MAX_WORLD];   // base of rank r's flag array
    unsigned int* done;                  // local CTA completion counter
    long long nq_max;
    int world, rank;
    unsigned int epoch;                  // 0: P2P disabled
    hamx_top2* merge_out;                // non-NULL: the CTA that publishes also waits for the world and merges (small nq)
    long long merge_nq;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void p2p_store(const P2PView& pv, long long qi, const hamx_top2& v)
{
    const size_t slot = ((size_t)(pv.epoch & 1u) * pv.world + pv.rank) * (size_t)pv.nq_max + (size_t)qi;
    const uint4 bits = make_uint4((uint32_t)v.dist0, (uint32_t)v.idx0, (uint32_t)v.dist1, (uint32_t)v.idx1);
    for (int r = 0; r < pv.world; r++) *reinterpret_cast<uint4*>(pv.gather[r] + slot) = bits;
}

__device__ __forceinline__ hamx_top2 p2p_merge_one(const P2PView& pv, long long i)
{
    const hamx_top2* base = pv.gather[pv.rank] + (size_t)(pv.epoch & 1u) * pv.world * (size_t)pv.nq_max;
    unsigned long long m0 = ~0ull, m1 = ~0ull;
    for (int r = 0; r < pv.world; r++) {
        const uint4 bits = __ldcg(reinterpret_cast<const uint4*>(base + (size_t)r * pv.nq_max + i));   // written by a peer: skip L1
        top2_insert64(m0, m1, top2_key64((int32_t)bits.x, (int32_t)bits.y));
        top2_insert64(m0, m1, top2_key64((int32_t)bits.z, (int32_t)bits.w));
    }
    hamx_top2 r;
    r.dist0 = m0 == ~0ull ? -1 : (int32_t)(m0 >> 32);
    r.idx0 = m0 == ~0ull ? -1 : (int32_t)(m0 & 0xFFFFFFFFu);
    r.dist1 = m1 == ~0ull ? -1 : (int32_t)(m1 >> 32);
    r.idx1 = m1 == ~0ull ? -1 : (int32_t)(m1 & 0xFFFFFFFFu);
    return r;
}

__device__ __forceinline__ void p2p_wait_world(const P2PView& pv)
{
    if (threadIdx.x < pv.world) {
        const unsigned int* f = pv.flags[pv.rank] + (pv.epoch & 1u) * pv.world + threadIdx.x;
        while ((int)(ld_acquire_sys(f) - pv.epoch) < 0) __nanosleep(200);
    }
    __syncthreads();
}

// called by every CTA that wrote results, after its last p2p_store; `writers` = number of such CTAs in the grid (one per
// block of HT_QB queries).  The last of them publishes this rank's epoch to the world.  With pv.merge_out set (small query
// sets) every writer then waits for the world's flags and merges ITS OWN block of queries, so one kernel launch per rank
// computes, exchanges and reduces -- and the reduction is spread over the writers instead of queueing behind one CTA (a
// single CTA merging 2000 queries x 8 ranks cost as much as the all-gather it replaces).  At most 64 writers spin (the host
// fuses only up to 16 k queries), far fewer than SMs, so the CTAs of the grid that have not run yet always find one.
__device__ __forceinline__ void p2p_publish(const P2PView& pv, unsigned int writers)
{
    __threadfence_system();     // this thread's peer stores are visible system-wide before the flag can be
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(pv.done, 1u);
        if (prev == writers - 1) {
            *pv.done = 0;       // ready for the next launch (stream-ordered)
            __threadfence_system();
            for (int r = 0; r < pv.world; r++) st_release_sys(pv.flags[r] + (pv.epoch & 1u) * pv.world + pv.rank, pv.epoch);
        }
    }
    if (pv.merge_out == nullptr) return;
    p2p_wait_world(pv);
    const long long q0 = (long long)blockIdx.x * HT_QB, q1 = q0 + HT_QB < pv.merge_nq ? q0 + HT_QB : pv.merge_nq;
    for (long long i = q0 + threadIdx.x; i < q1; i += blockDim.x) pv.merge_out[i] = p2p_merge_one(pv, i);
}

// grid = (query blocks, train splits, pairs).  partial is [pair][nsplit][nq_stride] (only touched when nsplit > 1); the
// last CTA of a query block to finish merges the splits, so one launch yields final results.  With `pairs` == NULL the
// launch handles the single problem `one`; otherwise pair blockIdx.z of a device-resident table (batched matching of
// many small frame pairs in one launch).
template <bool P2P>
__global__ void __launch_bounds__(HT_THREADS)
k_hamming_knn2(const hamx_pair one, const hamx_pair* __restrict__ pairs, int tiles_per_split, uint2* partial,
               size_t nq_stride, unsigned int* arrivals, int qblocks_stride, hamx_top2* __restrict__ out_base,
               size_t out_stride, int64_t idx_offset, const __grid_constant__ P2PView pv)
{
    __shared__ __align__(128) uint4 s_tile[HT_STAGES][HT_TT * 2];
    __shared__ __align__(8) uint64_t s_full[HT_STAGES];
    __shared__ int s_last;

    const hamx_pair pd = pairs ? pairs[blockIdx.z] : one;
    const int64_t nq = pd.nq;
    const int nt = pd.nt;
    if ((int64_t)blockIdx.x * HT_QB >= nq) return;   // batched launches are sized for the largest pair
    const uint4* __restrict__ q = reinterpret_cast<const uint4*>(pd.q);
    const uint4* __restrict__ t = reinterpret_cast<const uint4*>(pd.t);
    hamx_top2* __restrict__ out = out_base + (size_t)blockIdx.z * out_stride;
    partial += (size_t)blockIdx.z * gridDim.y * nq_stride;
    arrivals += (size_t)blockIdx.z * qblocks_stride;

    const int tid = threadIdx.x;
    const int nsplit = gridDim.y;
    const int ntiles_all = (nt + HT_TT - 1) / HT_TT;
    const int tile_begin = blockIdx.y * tiles_per_split;
    const int tile_end = min(tile_begin + tiles_per_split, ntiles_all);
    const int ntiles = max(tile_end - tile_begin, 0);

    if (tid == 0) {
        for (int s = 0; s < HT_STAGES; s++) mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int i) {   // thread 0 only: fetch tile (tile_begin + i) into stage i % STAGES
        int tile = tile_begin + i;
        int rows = min(HT_TT, nt - tile * HT_TT);
        uint32_t bytes = (uint32_t)rows * 32u;
        int st = i % HT_STAGES;
        mbar_expect_tx(&s_full[st], bytes);
        tma_bulk_g2s(&s_tile[st][0], t + (size_t)tile * HT_TT * 2, bytes, &s_full[st]);
    };
    if (tid == 0)
        for (int i = 0; i < HT_STAGES && i < ntiles; i++) issue(i);

    // this thread's queries: coalesced over the CTA (thread tid takes rows tid, tid + 128 of the block)
    uint32_t qr[HT_QPT][8];
    int64_t qi[HT_QPT];
    uint32_t b0[HT_QPT], b1[HT_QPT];
#pragma unroll
    for (int k = 0; k < HT_QPT; k++) {
        qi[k] = (int64_t)blockIdx.x * HT_QB + k * HT_THREADS + tid;
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (qi[k] < nq) { lo = q[2 * qi[k]]; hi = q[2 * qi[k] + 1]; }
        qr[k][0] = lo.x; qr[k][1] = lo.y; qr[k][2] = lo.z; qr[k][3] = lo.w;
        qr[k][4] = hi.x; qr[k][5] = hi.y; qr[k][6] = hi.z; qr[k][7] = hi.w;
        b0[k] = HT_NONE; b1[k] = HT_NONE;
    }

    for (int i = 0; i < ntiles; i++) {
        const int st = i % HT_STAGES;
        mbar_wait(&s_full[st], (uint32_t)(i / HT_STAGES) & 1u);
        const uint4* ts = &s_tile[st][0];
        const int tile = tile_begin + i;
        const uint32_t base = (uint32_t)tile * HT_TT;
        const int rows = min(HT_TT, nt - tile * HT_TT);
        if (rows == HT_TT) {
#pragma unroll 4
            for (int j = 0; j < HT_TT; j++) {
                uint4 a = ts[2 * j], b = ts[2 * j + 1];
#pragma unroll
                for (int k = 0; k < HT_QPT; k++) {
                    uint32_t d = ham256(qr[k], a, b);
                    top2_insert(b0[k], b1[k], (d << HT_IDX_BITS) + (base + j));
                }
            }
        } else {
            for (int j = 0; j < rows; j++) {
                uint4 a = ts[2 * j], b = ts[2 * j + 1];
#pragma unroll
                for (int k = 0; k < HT_QPT; k++) {
                    uint32_t d = ham256(qr[k], a, b);
                    top2_insert(b0[k], b1[k], (d << HT_IDX_BITS) + (base + j));
                }
            }
        }
        __syncthreads();   // everyone is done with stage st before the TMA engine overwrites it
        if (tid == 0 && i + HT_STAGES < ntiles) issue(i + HT_STAGES);
    }

    if (nsplit == 1) {
#pragma unroll
        for (int k = 0; k < HT_QPT; k++)
            if (qi[k] < nq) {
                const hamx_top2 v = decode_top2(b0[k], b1[k], idx_offset);
                if (P2P) p2p_store(pv, qi[k], v);
                else out[qi[k]] = v;
            }
        if (P2P) p2p_publish(pv, gridDim.x);
        return;
    }

#pragma unroll
    for (int k = 0; k < HT_QPT; k++)
        if (qi[k] < nq) __stcg(&partial[(size_t)blockIdx.y * nq_stride + qi[k]], make_uint2(b0[k], b1[k]));
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int prev = atomicAdd(&arrivals[blockIdx.x], 1u);
        s_last = (prev == (unsigned int)(nsplit - 1));
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int k = 0; k < HT_QPT; k++) {
        if (qi[k] >= nq) continue;
        uint32_t m0 = HT_NONE, m1 = HT_NONE;
        // the loads are independent (only the cheap min/max chain is serial): keep 8 in flight, or a merge over a few
        // hundred splits costs a few hundred L2 round trips at the tail of the kernel
        int s = 0;
        for (; s + 8 <= nsplit; s += 8) {
            uint2 p[8];
#pragma unroll
            for (int u = 0; u < 8; u++) p[u] = __ldcg(&partial[(size_t)(s + u) * nq_stride + qi[k]]);
#pragma unroll
            for (int u = 0; u < 8; u++) { top2_insert(m0, m1, p[u].x); top2_insert(m0, m1, p[u].y); }
        }
        for (; s < nsplit; s++) {
            uint2 p = __ldcg(&partial[(size_t)s * nq_stride + qi[k]]);
            top2_insert(m0, m1, p.x);
            top2_insert(m0, m1, p.y);
        }
        const hamx_top2 v = decode_top2(m0, m1, idx_offset);
        if (P2P) p2p_store(pv, qi[k], v);
        else out[qi[k]] = v;
    }
    if (tid == 0) arrivals[blockIdx.x] = 0;   // ready for the next launch
    if (P2P) p2p_publish(pv, gridDim.x);      // one merging CTA per query block
}

// Scatter of an already computed local result (train sets that needed several launches) with the same signalling.
__global__ void __launch_bounds__(256) k_scatter_p2p(const hamx_top2* __restrict__ local, long long nq, const __grid_constant__ P2PView pv)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (long long)gridDim.x * blockDim.x)
        p2p_store(pv, i, local[i]);
    p2p_publish(pv, gridDim.x);
}

// Waits until every rank has published `epoch`, then merges the world's slots of this rank's gather buffer.
__global__ void __launch_bounds__(256) k_merge_top2_p2p(long long nq, hamx_top2* __restrict__ out, const __grid_constant__ P2PView pv)
{
    p2p_wait_world(pv);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) out[i] = p2p_merge_one(pv, i);
}

// parts is [nparts][nq]; lexicographic (distance, index) merge, identical to a single-device run over the union.
__global__ void k_merge_top2(const hamx_top2* __restrict__ parts, int nparts, int64_t nq, hamx_top2* __restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    unsigned long long m0 = ~0ull, m1 = ~0ull;
    for (int p = 0; p < nparts; p++) {
        hamx_top2 v = parts[(size_t)p * nq + i];
        top2_insert64(m0, m1, top2_key64(v.dist0, v.idx0));
        top2_insert64(m0, m1, top2_key64(v.dist1, v.idx1));
    }
    hamx_top2 r;
    r.dist0 = m0 == ~0ull ? -1 : (int32_t)(m0 >> 32);
    r.idx0 = m0 == ~0ull ? -1 : (int32_t)(m0 & 0xFFFFFFFFu);
    r.dist1 = m1 == ~0ull ? -1 : (int32_t)(m1 >> 32);
    r.idx1 = m1 == ~0ull ? -1 : (int32_t)(m1 & 0xFFFFFFFFu);
    out[i] = r;
}

__global__ void k_top2_to_dmatch(const hamx_top2* __restrict__ top2, int64_t nq, orbx_dmatch* __restrict__ out,
                                 int32_t* __restrict__ counts)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    hamx_top2 v = top2[i];
    orbx_dmatch a = { (int32_t)i, v.idx0, 0, (float)v.dist0 };
    orbx_dmatch b = { (int32_t)i, v.idx1, 0, (float)v.dist1 };
    out[2 * i] = a;
    out[2 * i + 1] = b;
    counts[i] = (v.idx0 >= 0) + (v.idx1 >= 0);
}

// Lowe ratio (src/CameraPoseEstimator.cpp:208-212: float multiply, strict <) + order-preserving compaction.
// One CTA walks the queries in chunks of 1024 so the accepted list stays in ascending query order.
__global__ void __launch_bounds__(1024) k_ratio_compact(const hamx_top2* __restrict__ top2_base, size_t top2_stride, int64_t nq_one,
                                                        const hamx_pair* __restrict__ pairs, float ratio,
                                                        orbx_dmatch* __restrict__ good_base, size_t good_stride,
                                                        long long* __restrict__ ngood_base)
{
    const hamx_top2* __restrict__ top2 = top2_base + (size_t)blockIdx.x * top2_stride;
    orbx_dmatch* __restrict__ good = good_base + (size_t)blockIdx.x * good_stride;
    long long* __restrict__ ngood = ngood_base + blockIdx.x;
    const int64_t nq = pairs ? pairs[blockIdx.x].nq : nq_one;
    __shared__ int s_warp[32];
    __shared__ long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int64_t c = 0; c < nq; c += 1024) {
        int64_t i = c + tid;
        bool ok = false;
        hamx_top2 v = { 0, 0, 0, 0 };
        if (i < nq) {
            v = top2[i];
            ok = v.idx0 >= 0 && v.idx1 >= 0 && (float)v.dist0 < __fmul_rn((float)v.dist1, ratio);
        }
        unsigned int bal = __ballot_sync(0xFFFFFFFFu, ok);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; w++) {   // 32 broadcast reads; negligible next to the matching kernel
            int n = s_warp[w];
            before += w < wid ? n : 0;
            total += n;
        }
        long long base = s_base;
        if (ok) {
            long long pos = base + before + __popc(bal & ((1u << lane) - 1u));
            orbx_dmatch m = { (int32_t)i, v.idx0, 0, (float)v.dist0 };
            good[pos] = m;
        }
        __syncthreads();
        if (tid == 0) s_base = base + total;
        __syncthreads();
    }
    if (tid == 0) *ngood = s_base;
}

// Pair table for consecutive-frame matching of a batch: frame f (query) against frame f-1 (train); frame 0 against the
// last frame of the previous batch if there is one.  Counts live on the device, so no host round trip is needed.
__global__ void k_build_consecutive_pairs(const uint8_t* desc, const int32_t* counts, int nframes, int cap,
                                          const uint8_t* prev_desc, const int32_t* prev_count, hamx_pair* pairs)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    hamx_pair p;
    p.q = desc + (size_t)f * cap * 32;
    p.nq = min(counts[f], cap);
    if (f > 0) {
        p.t = desc + (size_t)(f - 1) * cap * 32;
        p.nt = min(counts[f - 1], cap);
    } else if (prev_desc) {
        p.t = prev_desc;
        p.nt = min(*prev_count, cap);
    } else {
        p.t = desc;
        p.nt = 0;
        p.nq = 0;   // the very first frame of a sequence has nothing to match against
    }
    pairs[f] = p;
}

// Pair table for the steady-state pattern of CameraPoseEstimator::pnpPoseEstimation (src/CameraPoseEstimator.cpp:405-409):
// frame f (query) against each of its `back` predecessors f-1 .. f-back (train).  Predecessors before the start of the
// batch come from the history (hist[0] = most recent frame before the batch); pair index = f * back + (j - 1).
__global__ void k_build_back_pairs(const uint8_t* desc, const int32_t* counts, int nframes, int cap, int back, const uint8_t* hist_desc,
                                   const int32_t* hist_counts, int nhist, hamx_pair* pairs)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nframes * back) return;
    const int f = idx / back, j = idx - f * back + 1;
    hamx_pair p;
    p.q = desc + (size_t)f * cap * 32;
    p.nq = min(counts[f], cap);
    const int src = f - j;
    if (src >= 0) {
        p.t = desc + (size_t)src * cap * 32;
        p.nt = min(counts[src], cap);
    } else if (-src - 1 < nhist) {
        p.t = hist_desc + (size_t)(-src - 1) * cap * 32;
        p.nt = min(hist_counts[-src - 1], cap);
    } else {
        p.t = desc;
        p.nt = 0;
        p.nq = 0;   // fewer than j frames exist before frame f
    }
    pairs[idx] = p;
}

// history' = the `back` most recent frames after the batch, most recent first (from the batch, then from the old history)
__global__ void k_update_history(const uint8_t* desc, const int32_t* counts, int nframes, int cap, const uint8_t* old_desc,
                                 const int32_t* old_counts, int nold, uint8_t* new_desc, int32_t* new_counts)
{
    const int hi = blockIdx.x;
    const int src = nframes - 1 - hi;
    const uint8_t* from = nullptr;
    int n = 0;
    if (src >= 0) { from = desc + (size_t)src * cap * 32; n = min(counts[src], cap); }
    else if (-src - 1 < nold) { from = old_desc + (size_t)(-src - 1) * cap * 32; n = min(old_counts[-src - 1], cap); }
    if (threadIdx.x == 0) new_counts[hi] = n;
    const uint4* a = reinterpret_cast<const uint4*>(from);
    uint4* b = reinterpret_cast<uint4*>(new_desc + (size_t)hi * cap * 32);
    for (int i = threadIdx.x; i < n * 2; i += blockDim.x) b[i] = a[i];
}

// ---- loop-closure candidate scoring (LoopCloser::DetectLoop / NBestMatches, reference src/LoopCloser.cpp:19-105) ---------
// Every thread walks the train rows [0, nt) of one set in ascending order through the same TMA ring as k_hamming_knn2 and
// calls row(j, a, b) for each (a, b = the two halves of train row j; the same row for all threads of the CTA).
template <class Row>
__device__ __forceinline__ void ham_stream_rows(const uint4* __restrict__ t, int nt, uint4 (*s_tile)[HT_TT * 2], uint64_t* s_full, Row row)
{
    const int tid = threadIdx.x;
    const int ntiles = (nt + HT_TT - 1) / HT_TT;
    if (tid == 0) {
        for (int s = 0; s < HT_STAGES; s++) mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int i) {
        const int rows = min(HT_TT, nt - i * HT_TT);
        const uint32_t bytes = (uint32_t)rows * 32u;
        const int st = i % HT_STAGES;
        mbar_expect_tx(&s_full[st], bytes);
        tma_bulk_g2s(&s_tile[st][0], t + (size_t)i * HT_TT * 2, bytes, &s_full[st]);
    };
    if (tid == 0)
        for (int i = 0; i < HT_STAGES && i < ntiles; i++) issue(i);
    for (int i = 0; i < ntiles; i++) {
        const int st = i % HT_STAGES;
        mbar_wait(&s_full[st], (uint32_t)(i / HT_STAGES) & 1u);
        const uint4* ts = &s_tile[st][0];
        const int rows = min(HT_TT, nt - i * HT_TT);
        if (rows == HT_TT) {
#pragma unroll 4
            for (int j = 0; j < HT_TT; j++) row(i * HT_TT + j, ts[2 * j], ts[2 * j + 1]);
        } else {
            for (int j = 0; j < rows; j++) row(i * HT_TT + j, ts[2 * j], ts[2 * j + 1]);
        }
        __syncthreads();
        if (tid == 0 && i + HT_STAGES < ntiles) issue(i + HT_STAGES);
    }
}

// DetectLoop's score of one stored frame (:34-41) is the number of n-best distances below the threshold, summed over the
// current frame's descriptors.  The n best distances of a query are its n smallest, so that number is
// min(n, #{train rows closer than thr}) -- no list has to be built (tests/test_oracle_tri.py checks the identity against the
// literal lists).  grid = (query blocks, stored frames); a CTA counts for 256 queries in registers, clips, reduces and adds
// its sum to the frame's score.
__global__ void __launch_bounds__(HT_THREADS)
k_loop_score(const uint8_t* __restrict__ q_, int nq, const uint8_t* __restrict__ frames, const int32_t* __restrict__ counts, int cap, int nbest,
             uint32_t thr, int32_t* __restrict__ scores)
{
    __shared__ __align__(128) uint4 s_tile[HT_STAGES][HT_TT * 2];
    __shared__ __align__(8) uint64_t s_full[HT_STAGES];
    __shared__ int s_sum;
    const int f = blockIdx.y, tid = threadIdx.x;
    const int nt = min(counts[f], cap);
    if (nt <= 0) return;
    if (tid == 0) s_sum = 0;
    const uint4* __restrict__ q = reinterpret_cast<const uint4*>(q_);
    uint32_t qr[HT_QPT][8];
    uint32_t cnt[HT_QPT];
    bool live[HT_QPT];
#pragma unroll
    for (int k = 0; k < HT_QPT; k++) {
        const int qi = blockIdx.x * HT_QB + k * HT_THREADS + tid;
        live[k] = qi < nq;
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (live[k]) { lo = q[2 * (size_t)qi]; hi = q[2 * (size_t)qi + 1]; }
        qr[k][0] = lo.x; qr[k][1] = lo.y; qr[k][2] = lo.z; qr[k][3] = lo.w;
        qr[k][4] = hi.x; qr[k][5] = hi.y; qr[k][6] = hi.z; qr[k][7] = hi.w;
        cnt[k] = 0;
    }
    ham_stream_rows(reinterpret_cast<const uint4*>(frames + (size_t)f * cap * 32), nt, s_tile, s_full, [&](int, const uint4& a, const uint4& b) {
#pragma unroll
        for (int k = 0; k < HT_QPT; k++) cnt[k] += ham256(qr[k], a, b) < thr ? 1u : 0u;
    });
    int sum = 0;
#pragma unroll
    for (int k = 0; k < HT_QPT; k++) sum += live[k] ? (int)min(cnt[k], (uint32_t)nbest) : 0;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if ((tid & 31) == 0 && sum) atomicAdd(&s_sum, sum);
    __syncthreads();
    if (tid == 0 && s_sum) atomicAdd(&scores[f], s_sum);
}

// :42-46 -- `if (count > count_max)` with count_max starting at 0: the first frame with the strictly largest non-zero score.
// best[0] = that frame (-1: every score is zero), best[1] = its score.
__global__ void __launch_bounds__(256) k_loop_best(const int32_t* __restrict__ scores, int nframes, int32_t* __restrict__ best)
{
    __shared__ unsigned long long s_key[256];
    // larger score wins, then the LOWER index: key = score << 32 | ~index, maximised
    unsigned long long key = 0;
    for (int i = threadIdx.x; i < nframes; i += blockDim.x) {
        const int sc = scores[i];
        if (sc > 0) {
            const unsigned long long k = ((unsigned long long)(uint32_t)sc << 32) | (uint32_t)~(uint32_t)i;
            key = k > key ? k : key;
        }
    }
    s_key[threadIdx.x] = key;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if (threadIdx.x < d) { const unsigned long long o = s_key[threadIdx.x + d]; if (o > s_key[threadIdx.x]) s_key[threadIdx.x] = o; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long k = s_key[0];
        best[0] = k ? (int32_t)~(uint32_t)(k & 0xFFFFFFFFu) : -1;
        best[1] = (int32_t)(k >> 32);
    }
}

// NBestMatches itself (:53-105): one thread per query walks the train rows in order and inserts each candidate exactly like
// the reference -- it takes the first slot whose distance it is strictly below and the displaced entry is carried down
// under the same strict test (:86-101).  Among equal distances that is not the lexicographic order of k_hamming_knn2 (a
// displaced entry skips its equals and may be the one that drops out), so the insertion is reproduced literally rather
// than through packed keys; a candidate that is not below the last slot cannot enter and is skipped.
template <int N>
__global__ void __launch_bounds__(HT_THREADS)
k_nbest(const uint8_t* __restrict__ q_, int nq, const uint8_t* __restrict__ t_, int nt, int n, int32_t* __restrict__ dist, int32_t* __restrict__ idx)
{
    __shared__ __align__(128) uint4 s_tile[HT_STAGES][HT_TT * 2];
    __shared__ __align__(8) uint64_t s_full[HT_STAGES];
    const int qi = blockIdx.x * HT_THREADS + threadIdx.x;
    const uint4* __restrict__ q = reinterpret_cast<const uint4*>(q_);
    uint32_t qr[8];
    {
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (qi < nq) { lo = q[2 * (size_t)qi]; hi = q[2 * (size_t)qi + 1]; }
        qr[0] = lo.x; qr[1] = lo.y; qr[2] = lo.z; qr[3] = lo.w; qr[4] = hi.x; qr[5] = hi.y; qr[6] = hi.z; qr[7] = hi.w;
    }
    int32_t cd[N], ci[N];
#pragma unroll
    for (int k = 0; k < N; k++) { cd[k] = 0x7FFFFFFF; ci[k] = -1; }
    ham_stream_rows(reinterpret_cast<const uint4*>(t_), nt, s_tile, s_full, [&](int j, const uint4& a, const uint4& b) {
        int32_t d = (int32_t)ham256(qr, a, b), c = j;
        if (d < cd[N - 1]) {
#pragma unroll
            for (int k = 0; k < N; k++) {
                const bool lt = d < cd[k];
                const int32_t od = cd[k], oi = ci[k];
                cd[k] = lt ? d : od; ci[k] = lt ? c : oi;
                d = lt ? od : d; c = lt ? oi : c;
            }
        }
    });
    if (qi >= nq) return;
#pragma unroll
    for (int k = 0; k < N; k++)
        if (k < n) {
            dist[(size_t)qi * n + k] = ci[k] < 0 ? -1 : cd[k];
            idx[(size_t)qi * n + k] = ci[k];
        }
}

// Register-only POPC throughput probe: 8 independent POPC + 8 XOR per iteration and thread, like the matcher's inner
// loop without its memory traffic.
__global__ void __launch_bounds__(1024) k_popc_peak(uint32_t* sink, int iters, uint32_t seed)
{
    uint32_t x0 = seed ^ threadIdx.x, x1 = x0 * 2654435761u, x2 = x1 ^ 0x9E3779B9u, x3 = x2 * 40503u;
    uint32_t x4 = ~x0, x5 = ~x1, x6 = ~x2, x7 = ~x3;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        uint32_t m = (uint32_t)i * 0x01010101u;
        a0 += __popc(x0 ^ m) + __popc(x1 ^ m);
        a1 += __popc(x2 ^ m) + __popc(x3 ^ m);
        a2 += __popc(x4 ^ m) + __popc(x5 ^ m);
        a3 += __popc(x6 ^ m) + __popc(x7 ^ m);
    }
    uint32_t a = a0 + a1 + a2 + a3;
    if (a == 0xFFFFFFFFu) sink[0] = a;   // never true; keeps the loop alive
}

}  // namespace
}  // namespace orbx

#include "hamming_tc.cuh"

// ------------------------------------------------------------------------------------------------ host side
using namespace orbx;

struct hamx_context {
    int device;
    cudaStream_t own_stream, stream;
    // workspaces (grown on demand)
    uint8_t* d_q; size_t q_bytes;
    uint8_t* d_t; size_t t_bytes;
    uint2* d_partial; size_t partial_bytes;
    unsigned int* d_arrivals; size_t arrivals_n;
    hamx_top2* d_top2; size_t top2_bytes;
    hamx_top2* d_parts; size_t parts_bytes;
    orbx_dmatch* d_dm; size_t dm_bytes;
    int32_t* d_counts; size_t counts_bytes;
    long long* d_ngood;
    hamx_pair* d_pairs; size_t pairs_bytes;
    int sm_count;
    // peer-memory path (hamx_p2p_*)
    P2PView pv;                        // host copy; epoch advances per collective call
    void* p2p_buf;                     // this rank's exported allocation: flags, then the gather buffer
    size_t p2p_bytes;
    void* p2p_opened[HT_MAX_WORLD];    // peer mappings opened through cudaIpc (closed by hamx_p2p_close)
    hamx_top2* d_p2p_local; size_t p2p_local_bytes;
    // tensor-core path: the train set expanded to +-1 bytes in the MMA operand image (hamming_tc.cuh)
    uint8_t* d_texp; size_t texp_bytes;
    bool tc_ready, loop_tc_ready;
    int kernel_mode;                   // HAMX_KERNEL_*
};

static const P2PView kNoP2P = {};

template <typename T>
static int grow(T** p, size_t* have, size_t want, bool zero = false, cudaStream_t s = 0)
{
    if (*have >= want && *p) return ORBX_OK;
    if (*p) { cudaError_t e = cudaFree(*p); (void)e; *p = nullptr; *have = 0; }
    size_t bytes = align_up(want + want / 4, 256);
    cudaError_t e = cudaMalloc((void**)p, bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return ORBX_E_ALLOC; }
    if (zero) { e = cudaMemsetAsync(*p, 0, bytes, s); if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return ORBX_E_CUDA; } }
    *have = bytes;
    return ORBX_OK;
}

extern "C" int hamx_create(hamx_handle* out, int device)
{
    ORBX_REQUIRE(out != nullptr, "hamx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { set_error("hamx_create: no CUDA device (%s); liborbx has no CPU fallback", cudaGetErrorString(e)); return ORBX_E_CUDA; }
    ORBX_REQUIRE(device >= 0 && device < ndev, "hamx_create: device %d out of range [0,%d)", device, ndev);
    ORBX_CUDA(cudaSetDevice(device));
    hamx_context* h = new hamx_context();
    memset(h, 0, sizeof(*h));
    h->device = device;
    ORBX_CUDA_OR(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking), hamx_destroy(h));
    h->stream = h->own_stream;
    ORBX_CUDA_OR(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device), hamx_destroy(h));
    ORBX_CUDA_OR(cudaMalloc((void**)&h->d_ngood, sizeof(long long)), hamx_destroy(h));
    *out = h;
    return ORBX_OK;
}

extern "C" int hamx_destroy(hamx_handle h)
{
    if (!h) return ORBX_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_q); cudaFree(h->d_t); cudaFree(h->d_partial); cudaFree(h->d_arrivals); cudaFree(h->d_top2);
    cudaFree(h->d_parts); cudaFree(h->d_dm); cudaFree(h->d_counts); cudaFree(h->d_ngood); cudaFree(h->d_pairs);
    hamx_p2p_close(h);
    cudaFree(h->d_p2p_local);
    cudaFree(h->d_texp);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return ORBX_OK;
}

extern "C" int hamx_set_stream(hamx_handle h, void* cuda_stream)
{
    ORBX_REQUIRE(h != nullptr, "hamx_set_stream: NULL handle");
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return ORBX_OK;
}

extern "C" int hamx_set_kernel(hamx_handle h, int mode)
{
    ORBX_REQUIRE(h != nullptr, "hamx_set_kernel: NULL handle");
    ORBX_REQUIRE(mode == HAMX_KERNEL_AUTO || mode == HAMX_KERNEL_INTEGER || mode == HAMX_KERNEL_TENSOR, "hamx_set_kernel: unknown mode %d", mode);
    h->kernel_mode = mode;
    return ORBX_OK;
}

extern "C" int hamx_get_stream(hamx_handle h, void** cuda_stream)
{
    ORBX_REQUIRE(h != nullptr && cuda_stream != nullptr, "hamx_get_stream: NULL argument");
    *cuda_stream = (void*)h->stream;
    return ORBX_OK;
}

extern "C" int hamx_synchronize(hamx_handle h)
{
    ORBX_REQUIRE(h != nullptr, "hamx_synchronize: NULL handle");
    ORBX_CUDA(cudaSetDevice(h->device));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

static int fill_absent(hamx_handle h, hamx_top2* d_out, int64_t nq)
{
    // nt == 0: every query has no neighbours (all fields -1 == 0xFF bytes)
    ORBX_CUDA(cudaMemsetAsync(d_out, 0xFF, (size_t)nq * sizeof(hamx_top2), h->stream));
    return ORBX_OK;
}

// CTAs wanted per SM before the train range stops being split (measured on B200, 2000 x 2000 pairs and 2000 x 200 k: 24)
constexpr int HT_SPLIT_MULT = 24;

static void plan_split(int sm_count, int64_t nqb, int ntiles, int64_t npairs, int* nsplit_out, int* tps_out)
{
    const int64_t target = (int64_t)sm_count * HT_SPLIT_MULT;
    int64_t nsplit = nqb * npairs >= target ? 1 : (target + nqb * npairs - 1) / (nqb * npairs);
    if (nsplit > ntiles) nsplit = ntiles;
    if (nsplit > 65535) nsplit = 65535;
    if (nsplit < 1) nsplit = 1;
    int tps = (int)((ntiles + nsplit - 1) / nsplit);
    if (tps < 1) tps = 1;
    nsplit = (ntiles + tps - 1) / tps;
    if (nsplit < 1) nsplit = 1;
    *nsplit_out = (int)nsplit;
    *tps_out = tps;
}

// ---- tensor-core path (hamming_tc.cuh): the same contract as launch_chunk's integer-pipe kernel
static int tc_prepare(hamx_handle h)
{
    if (h->tc_ready) return ORBX_OK;
    ORBX_CUDA(cudaFuncSetAttribute(k_hamming_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    ORBX_CUDA(cudaFuncSetAttribute(k_hamming_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    h->tc_ready = true;
    return ORBX_OK;
}

static int launch_chunk_tc(hamx_handle h, const uint8_t* d_q, int64_t nq, const uint8_t* d_t, int nt, int64_t offset, hamx_top2* d_out,
                           const P2PView* pv)
{
    int rc = tc_prepare(h);
    if (rc) return rc;
    const int ntiles = (nt + TC_TN - 1) / TC_TN;
    rc = grow(&h->d_texp, &h->texp_bytes, (size_t)ntiles * TC_TILE_BYTES);
    if (rc) return rc;
    const long long units = (long long)ntiles * (TC_TILE_BYTES / 16);
    k_expand_train<<<(unsigned int)((units + 255) / 256), 256, 0, h->stream>>>(d_t, nt, nullptr, 0, reinterpret_cast<uint4*>(h->d_texp));
    ORBX_CUDA(cudaGetLastError());
    const int64_t nqb = (nq + TC_QB - 1) / TC_QB;
    // one CTA per SM: split the train range until every SM has a CTA (a few, for balance, when there are few query blocks)
    int nsplit = nqb >= 2 * h->sm_count ? 1 : (int)std::min<int64_t>((2 * h->sm_count + nqb - 1) / nqb, ntiles);
    if (nsplit > 65535) nsplit = 65535;
    const int tps = (ntiles + nsplit - 1) / nsplit;
    nsplit = (ntiles + tps - 1) / tps;
    if (nsplit > 1) {
        rc = grow(&h->d_partial, &h->partial_bytes, (size_t)nsplit * nq * sizeof(uint2));
        if (rc) return rc;
        rc = grow(&h->d_arrivals, &h->arrivals_n, (size_t)nqb * sizeof(unsigned int), true, h->stream);
        if (rc) return rc;
    }
    const dim3 grid((unsigned int)nqb, (unsigned int)nsplit, 1);
    if (pv)
        k_hamming_tc<true><<<grid, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(d_q, nq, h->d_texp, nt, nullptr, 0, tps, h->d_partial, (size_t)nq,
                                                                           h->d_arrivals, (int)nqb, d_out, 0, offset, *pv);
    else
        k_hamming_tc<false><<<grid, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(d_q, nq, h->d_texp, nt, nullptr, 0, tps, h->d_partial, (size_t)nq,
                                                                            h->d_arrivals, (int)nqb, d_out, 0, offset, kNoP2P);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

// Which kernel computes a single (query set, train set) problem.  The tensor-core kernel pays a fixed price per CTA (TMEM
// allocation, expanding 256 queries, the train set expanded once per call); below ~4 M comparisons the integer-pipe kernel
// is as fast or faster.
static bool use_tensor_cores(hamx_handle h, int64_t nq, int nt)
{
    if (h->kernel_mode == HAMX_KERNEL_INTEGER) return false;
    if (h->kernel_mode == HAMX_KERNEL_TENSOR) return true;
    return nq * (int64_t)nt >= (1ll << 22);
}

static int launch_chunk(hamx_handle h, const uint8_t* d_q, int64_t nq, const uint8_t* d_t, int nt, int64_t offset, hamx_top2* d_out,
                        const P2PView* pv = nullptr)
{
    if (use_tensor_cores(h, nq, nt)) return launch_chunk_tc(h, d_q, nq, d_t, nt, offset, d_out, pv);
    const int64_t nqb = (nq + HT_QB - 1) / HT_QB;
    const int ntiles = (nt + HT_TT - 1) / HT_TT;
    int nsplit, tps;
    plan_split(h->sm_count, nqb, ntiles, 1, &nsplit, &tps);
    if (nsplit > 1) {
        int rc = grow(&h->d_partial, &h->partial_bytes, (size_t)nsplit * nq * sizeof(uint2));
        if (rc) return rc;
        rc = grow(&h->d_arrivals, &h->arrivals_n, (size_t)nqb * sizeof(unsigned int), true, h->stream);
        if (rc) return rc;
    }
    hamx_pair one;
    one.q = d_q; one.t = d_t; one.nq = (int32_t)nq; one.nt = nt;
    dim3 grid((unsigned int)nqb, (unsigned int)nsplit, 1);
    if (pv)
        k_hamming_knn2<true><<<grid, HT_THREADS, 0, h->stream>>>(one, nullptr, tps, h->d_partial, (size_t)nq, h->d_arrivals, (int)nqb,
                                                                  d_out, 0, offset, *pv);
    else
        k_hamming_knn2<false><<<grid, HT_THREADS, 0, h->stream>>>(one, nullptr, tps, h->d_partial, (size_t)nq, h->d_arrivals, (int)nqb,
                                                                   d_out, 0, offset, kNoP2P);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

// Pre-size every workspace the _dev entry points can need for problems of up to nq queries, nt train rows and npairs
// batched pairs, so that none of them allocates (cudaFree / cudaMalloc synchronise the whole device) once a pipeline runs.
extern "C" int hamx_reserve(hamx_handle h, int64_t nq, int64_t nt, int npairs)
{
    ORBX_REQUIRE(h != nullptr, "hamx_reserve: NULL handle");
    ORBX_REQUIRE(nq >= 0 && nq < (1ll << 31) && nt >= 0 && npairs >= 0, "hamx_reserve: bad sizes");
    ORBX_CUDA(cudaSetDevice(h->device));
    const int64_t np = std::max(npairs, 1);
    const int64_t nqb = (nq + HT_QB - 1) / HT_QB;
    // plan_split: nsplit <= target / (nqb * np) + 1, hence np * nsplit * nq <= target * HT_QB + np * nq for every smaller problem too
    const size_t partial = ((size_t)h->sm_count * HT_SPLIT_MULT * HT_QB + (size_t)np * (size_t)nq) * sizeof(uint2);
    int rc = grow(&h->d_partial, &h->partial_bytes, partial);
    if (!rc) rc = grow(&h->d_arrivals, &h->arrivals_n, (size_t)np * std::max<int64_t>(nqb, 1) * sizeof(unsigned int), true, h->stream);
    if (!rc) rc = grow(&h->d_top2, &h->top2_bytes, (size_t)np * nq * sizeof(hamx_top2) + 16);
    if (!rc) rc = grow(&h->d_pairs, &h->pairs_bytes, (size_t)np * sizeof(hamx_pair));
    const int64_t chunk = 1ll << HT_IDX_BITS;
    if (!rc && nt > chunk) rc = grow(&h->d_parts, &h->parts_bytes, (size_t)((nt + chunk - 1) / chunk) * nq * sizeof(hamx_top2));
    if (!rc && h->p2p_buf) rc = grow(&h->d_p2p_local, &h->p2p_local_bytes, (size_t)nq * sizeof(hamx_top2));
    if (!rc && h->kernel_mode != HAMX_KERNEL_INTEGER) {      // tensor-core path: expanded train set(s), and its own split plan
        const int64_t tiles = (std::min<int64_t>(nt, chunk) + TC_TN - 1) / TC_TN;
        rc = grow(&h->d_texp, &h->texp_bytes, (size_t)np * (size_t)std::max<int64_t>(tiles, 1) * TC_TILE_BYTES);
        if (!rc) rc = grow(&h->d_partial, &h->partial_bytes, std::max(partial, ((size_t)2 * h->sm_count * TC_QB + (size_t)np * (size_t)nq) * sizeof(uint2)));
    }
    if (rc) return rc;
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

// Many small (query, train) pairs in one launch: consecutive-frame matching of a whole batch of frames.
extern "C" int hamx_match_pairs_dev(hamx_handle h, const hamx_pair* d_pairs, int npairs, int max_nq, int max_nt, float ratio,
                                    orbx_dmatch* d_good, size_t good_stride, int64_t* d_ngood)
{
    ORBX_REQUIRE(h != nullptr, "hamx_match_pairs_dev: NULL handle");
    ORBX_REQUIRE(npairs >= 0 && max_nq >= 0 && max_nt >= 0 && max_nt <= (1 << HT_IDX_BITS), "hamx_match_pairs_dev: bad sizes");
    if (npairs == 0) return ORBX_OK;
    ORBX_REQUIRE(d_pairs && d_good && d_ngood && good_stride >= (size_t)max_nq, "hamx_match_pairs_dev: bad pointers or stride");
    ORBX_CUDA(cudaSetDevice(h->device));
    if (max_nq == 0) { ORBX_CUDA(cudaMemsetAsync(d_ngood, 0, (size_t)npairs * sizeof(int64_t), h->stream)); return ORBX_OK; }
    const int64_t nqb = (max_nq + HT_QB - 1) / HT_QB;
    int rc = grow(&h->d_top2, &h->top2_bytes, (size_t)npairs * max_nq * sizeof(hamx_top2) + 16);
    if (rc) return rc;
    if (h->kernel_mode != HAMX_KERNEL_INTEGER && max_nt > 0 && (h->kernel_mode == HAMX_KERNEL_TENSOR || (int64_t)max_nq * max_nt >= (1ll << 20))) {
        // tensor-core kernel, one launch for all pairs: every pair's train set is expanded into its own operand image first
        rc = tc_prepare(h);
        if (rc) return rc;
        const int tiles = (max_nt + TC_TN - 1) / TC_TN;
        const size_t pair_bytes = (size_t)tiles * TC_TILE_BYTES;
        rc = grow(&h->d_texp, &h->texp_bytes, (size_t)npairs * pair_bytes);
        if (rc) return rc;
        const long long units = (long long)tiles * (TC_TILE_BYTES / 16);
        k_expand_train<<<dim3((unsigned int)((units + 255) / 256), (unsigned int)npairs), 256, 0, h->stream>>>(nullptr, 0, d_pairs, pair_bytes / 16,
                                                                                                          reinterpret_cast<uint4*>(h->d_texp));
        ORBX_CUDA(cudaGetLastError());
        int nsplit = nqb * npairs >= 2 * h->sm_count ? 1 : (int)std::min<int64_t>((2 * h->sm_count + nqb * npairs - 1) / (nqb * npairs), tiles);
        const int tps = (tiles + nsplit - 1) / nsplit;
        nsplit = (tiles + tps - 1) / tps;
        if (nsplit > 1) {
            rc = grow(&h->d_partial, &h->partial_bytes, (size_t)npairs * nsplit * max_nq * sizeof(uint2));
            if (rc) return rc;
            rc = grow(&h->d_arrivals, &h->arrivals_n, (size_t)npairs * nqb * sizeof(unsigned int), true, h->stream);
            if (rc) return rc;
        }
        const dim3 grid((unsigned int)nqb, (unsigned int)nsplit, (unsigned int)npairs);
        k_hamming_tc<false><<<grid, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(nullptr, 0, h->d_texp, 0, d_pairs, pair_bytes, tps, h->d_partial,
                                                                            (size_t)max_nq, h->d_arrivals, (int)nqb, h->d_top2, (size_t)max_nq, 0,
                                                                            kNoP2P);
        ORBX_CUDA(cudaGetLastError());
        k_ratio_compact<<<npairs, 1024, 0, h->stream>>>(h->d_top2, (size_t)max_nq, 0, d_pairs, ratio, d_good, good_stride, (long long*)d_ngood);
        ORBX_CUDA(cudaGetLastError());
        return ORBX_OK;
    }
    const int ntiles = max_nt > 0 ? (max_nt + HT_TT - 1) / HT_TT : 1;
    int nsplit, tps;
    plan_split(h->sm_count, nqb, ntiles, npairs, &nsplit, &tps);
    if (nsplit > 1) {
        rc = grow(&h->d_partial, &h->partial_bytes, (size_t)npairs * nsplit * max_nq * sizeof(uint2));
        if (rc) return rc;
        rc = grow(&h->d_arrivals, &h->arrivals_n, (size_t)npairs * nqb * sizeof(unsigned int), true, h->stream);
        if (rc) return rc;
    }
    hamx_pair none;
    memset(&none, 0, sizeof(none));
    dim3 grid((unsigned int)nqb, (unsigned int)nsplit, (unsigned int)npairs);
    k_hamming_knn2<false><<<grid, HT_THREADS, 0, h->stream>>>(none, d_pairs, tps, h->d_partial, (size_t)max_nq, h->d_arrivals, (int)nqb,
                                                               h->d_top2, (size_t)max_nq, 0, kNoP2P);
    ORBX_CUDA(cudaGetLastError());
    k_ratio_compact<<<npairs, 1024, 0, h->stream>>>(h->d_top2, (size_t)max_nq, 0, d_pairs, ratio, d_good, good_stride, (long long*)d_ngood);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int hamx_match_consecutive_dev(hamx_handle h, const uint8_t* d_desc, const int32_t* d_counts, int nframes, int cap,
                                          const uint8_t* d_prev_desc, const int32_t* d_prev_count, float ratio,
                                          orbx_dmatch* d_good, int64_t* d_ngood)
{
    ORBX_REQUIRE(h != nullptr, "hamx_match_consecutive_dev: NULL handle");
    ORBX_REQUIRE(nframes >= 0 && cap >= 1, "hamx_match_consecutive_dev: bad sizes");
    if (nframes == 0) return ORBX_OK;
    ORBX_REQUIRE(d_desc && d_counts && d_good && d_ngood && (!d_prev_desc || d_prev_count), "hamx_match_consecutive_dev: NULL pointer");
    if ((((uintptr_t)d_desc) | ((uintptr_t)d_prev_desc)) & 15) { set_error("hamx_match_consecutive_dev: descriptor pointers must be 16-byte aligned"); return ORBX_E_ALIGN; }
    ORBX_CUDA(cudaSetDevice(h->device));
    int rc = grow(&h->d_pairs, &h->pairs_bytes, (size_t)nframes * sizeof(hamx_pair));
    if (rc) return rc;
    k_build_consecutive_pairs<<<(nframes + 127) / 128, 128, 0, h->stream>>>(d_desc, d_counts, nframes, cap, d_prev_desc, d_prev_count,
                                                                            h->d_pairs);
    ORBX_CUDA(cudaGetLastError());
    return hamx_match_pairs_dev(h, h->d_pairs, nframes, cap, cap, ratio, d_good, (size_t)cap, d_ngood);
}

extern "C" int hamx_match_back_dev(hamx_handle h, const uint8_t* d_desc, const int32_t* d_counts, int nframes, int cap, int back,
                                   const uint8_t* d_hist_desc, const int32_t* d_hist_counts, int nhist, float ratio, orbx_dmatch* d_good,
                                   int64_t* d_ngood)
{
    ORBX_REQUIRE(h != nullptr, "hamx_match_back_dev: NULL handle");
    ORBX_REQUIRE(nframes >= 0 && cap >= 1 && back >= 1 && back <= 64 && nhist >= 0, "hamx_match_back_dev: bad sizes");
    if (nframes == 0) return ORBX_OK;
    ORBX_REQUIRE(d_desc && d_counts && d_good && d_ngood && (nhist == 0 || (d_hist_desc && d_hist_counts)), "hamx_match_back_dev: NULL pointer");
    if ((((uintptr_t)d_desc) | ((uintptr_t)d_hist_desc)) & 15) { set_error("hamx_match_back_dev: descriptor pointers must be 16-byte aligned"); return ORBX_E_ALIGN; }
    ORBX_CUDA(cudaSetDevice(h->device));
    const int npairs = nframes * back;
    int rc = grow(&h->d_pairs, &h->pairs_bytes, (size_t)npairs * sizeof(hamx_pair));
    if (rc) return rc;
    k_build_back_pairs<<<(npairs + 127) / 128, 128, 0, h->stream>>>(d_desc, d_counts, nframes, cap, back, d_hist_desc, d_hist_counts, nhist, h->d_pairs);
    ORBX_CUDA(cudaGetLastError());
    return hamx_match_pairs_dev(h, h->d_pairs, npairs, cap, cap, ratio, d_good, (size_t)cap, d_ngood);
}

extern "C" int hamx_update_history_dev(hamx_handle h, const uint8_t* d_desc, const int32_t* d_counts, int nframes, int cap, int back,
                                       const uint8_t* d_old_desc, const int32_t* d_old_counts, int nold, uint8_t* d_new_desc,
                                       int32_t* d_new_counts)
{
    ORBX_REQUIRE(h != nullptr, "hamx_update_history_dev: NULL handle");
    ORBX_REQUIRE(nframes >= 0 && cap >= 1 && back >= 1 && nold >= 0 && d_new_desc && d_new_counts, "hamx_update_history_dev: bad arguments");
    ORBX_REQUIRE(d_new_desc != d_old_desc, "hamx_update_history_dev: the new history must not alias the old one");
    ORBX_CUDA(cudaSetDevice(h->device));
    k_update_history<<<back, 256, 0, h->stream>>>(d_desc, d_counts, nframes, cap, d_old_desc, d_old_counts, nold, d_new_desc, d_new_counts);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int hamx_knn2_dev(hamx_handle h, const uint8_t* d_q, int64_t nq, const uint8_t* d_t, int64_t nt, int64_t train_offset,
                             hamx_top2* d_out)
{
    ORBX_REQUIRE(h != nullptr, "hamx_knn2_dev: NULL handle");
    ORBX_REQUIRE(nq >= 0 && nt >= 0, "hamx_knn2_dev: negative size");
    ORBX_REQUIRE(nq < (1ll << 31) && nt + train_offset < (1ll << 31) && train_offset >= 0,
                 "hamx_knn2_dev: indices must fit int32 (cv::DMatch), got nq=%lld nt=%lld offset=%lld", (long long)nq,
                 (long long)nt, (long long)train_offset);
    if (nq == 0) return ORBX_OK;
    ORBX_REQUIRE(d_q && d_out && (nt == 0 || d_t), "hamx_knn2_dev: NULL pointer");
    if ((((uintptr_t)d_q) | ((uintptr_t)d_t) | ((uintptr_t)d_out)) & 15) { set_error("hamx_knn2_dev: device pointers must be 16-byte aligned"); return ORBX_E_ALIGN; }
    ORBX_CUDA(cudaSetDevice(h->device));
    if (nt == 0) return fill_absent(h, d_out, nq);
    const int64_t chunk = 1ll << HT_IDX_BITS;
    if (nt <= chunk) return launch_chunk(h, d_q, nq, d_t, (int)nt, train_offset, d_out);
    // > 2^23 train rows: per-chunk results merged with the 64-bit rule
    int nchunks = (int)((nt + chunk - 1) / chunk);
    int rc = grow(&h->d_parts, &h->parts_bytes, (size_t)nchunks * nq * sizeof(hamx_top2));
    if (rc) return rc;
    for (int c = 0; c < nchunks; c++) {
        int64_t lo = (int64_t)c * chunk, n = nt - lo < chunk ? nt - lo : chunk;
        rc = launch_chunk(h, d_q, nq, d_t + lo * 32, (int)n, train_offset + lo, h->d_parts + (size_t)c * nq);
        if (rc) return rc;
    }
    return hamx_merge_top2_dev(h, h->d_parts, nchunks, nq, d_out);
}

extern "C" int hamx_merge_top2_dev(hamx_handle h, const hamx_top2* d_parts, int nparts, int64_t nq, hamx_top2* d_out)
{
    ORBX_REQUIRE(h != nullptr, "hamx_merge_top2_dev: NULL handle");
    ORBX_REQUIRE(nparts >= 1 && nq >= 0, "hamx_merge_top2_dev: bad sizes");
    if (nq == 0) return ORBX_OK;
    ORBX_REQUIRE(d_parts && d_out, "hamx_merge_top2_dev: NULL pointer");
    ORBX_CUDA(cudaSetDevice(h->device));
    k_merge_top2<<<(unsigned int)((nq + 255) / 256), 256, 0, h->stream>>>(d_parts, nparts, nq, d_out);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int hamx_ratio_dev(hamx_handle h, const hamx_top2* d_top2, int64_t nq, float ratio, orbx_dmatch* d_good, int64_t* d_ngood)
{
    ORBX_REQUIRE(h != nullptr, "hamx_ratio_dev: NULL handle");
    ORBX_REQUIRE(nq >= 0 && d_ngood, "hamx_ratio_dev: bad arguments");
    ORBX_REQUIRE(nq == 0 || (d_top2 && d_good), "hamx_ratio_dev: NULL pointer");
    ORBX_CUDA(cudaSetDevice(h->device));
    k_ratio_compact<<<1, 1024, 0, h->stream>>>(d_top2, 0, nq, nullptr, ratio, d_good, 0, (long long*)d_ngood);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

static int upload_sets(hamx_handle h, const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt)
{
    int rc = grow(&h->d_q, &h->q_bytes, (size_t)nq * 32 + 32);
    if (rc) return rc;
    rc = grow(&h->d_t, &h->t_bytes, (size_t)nt * 32 + 32);
    if (rc) return rc;
    rc = grow(&h->d_top2, &h->top2_bytes, (size_t)nq * sizeof(hamx_top2) + 16);
    if (rc) return rc;
    if (nq) ORBX_CUDA(cudaMemcpyAsync(h->d_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, h->stream));
    if (nt) ORBX_CUDA(cudaMemcpyAsync(h->d_t, t, (size_t)nt * 32, cudaMemcpyHostToDevice, h->stream));
    return ORBX_OK;
}

extern "C" int hamx_knn2(hamx_handle h, const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, orbx_dmatch* out, int32_t* out_counts)
{
    ORBX_REQUIRE(h != nullptr, "hamx_knn2: NULL handle");
    ORBX_REQUIRE(nq >= 0 && nt >= 0, "hamx_knn2: negative size");
    if (nq == 0) return ORBX_OK;
    ORBX_REQUIRE(q && out && out_counts && (nt == 0 || t), "hamx_knn2: NULL pointer");
    ORBX_CUDA(cudaSetDevice(h->device));
    int rc = upload_sets(h, q, nq, t, nt);
    if (rc) return rc;
    rc = hamx_knn2_dev(h, h->d_q, nq, h->d_t, nt, 0, h->d_top2);
    if (rc) return rc;
    rc = grow(&h->d_dm, &h->dm_bytes, (size_t)nq * 2 * sizeof(orbx_dmatch));
    if (rc) return rc;
    rc = grow(&h->d_counts, &h->counts_bytes, (size_t)nq * sizeof(int32_t));
    if (rc) return rc;
    k_top2_to_dmatch<<<(unsigned int)((nq + 255) / 256), 256, 0, h->stream>>>(h->d_top2, nq, h->d_dm, h->d_counts);
    ORBX_CUDA(cudaGetLastError());
    ORBX_CUDA(cudaMemcpyAsync(out, h->d_dm, (size_t)nq * 2 * sizeof(orbx_dmatch), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(out_counts, h->d_counts, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int hamx_match_ratio(hamx_handle h, const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, float ratio,
                                orbx_dmatch* good, int64_t* ngood)
{
    ORBX_REQUIRE(h != nullptr, "hamx_match_ratio: NULL handle");
    ORBX_REQUIRE(nq >= 0 && nt >= 0 && ngood, "hamx_match_ratio: bad arguments");
    *ngood = 0;
    if (nq == 0) return ORBX_OK;
    ORBX_REQUIRE(q && good && (nt == 0 || t), "hamx_match_ratio: NULL pointer");
    ORBX_CUDA(cudaSetDevice(h->device));
    int rc = upload_sets(h, q, nq, t, nt);
    if (rc) return rc;
    rc = hamx_knn2_dev(h, h->d_q, nq, h->d_t, nt, 0, h->d_top2);
    if (rc) return rc;
    rc = grow(&h->d_dm, &h->dm_bytes, (size_t)nq * 2 * sizeof(orbx_dmatch));
    if (rc) return rc;
    rc = hamx_ratio_dev(h, h->d_top2, nq, ratio, h->d_dm, (int64_t*)h->d_ngood);
    if (rc) return rc;
    long long n = 0;
    ORBX_CUDA(cudaMemcpyAsync(&n, h->d_ngood, sizeof(n), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    if (n) ORBX_CUDA(cudaMemcpyAsync(good, h->d_dm, (size_t)n * sizeof(orbx_dmatch), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    *ngood = n;
    return ORBX_OK;
}

// ------------------------------------------------------------------------------------------------ loop-closure scoring
extern "C" int hamx_loop_score_dev(hamx_handle h, const uint8_t* d_q, int nq, const uint8_t* d_frames, const int32_t* d_counts, int nframes, int cap,
                                   int n, int thr, int32_t* d_scores, int32_t* d_best)
{
    ORBX_REQUIRE(h != nullptr, "hamx_loop_score_dev: NULL handle");
    ORBX_REQUIRE(nq >= 0 && nframes >= 0 && nframes <= 65535 && cap >= 1 && n >= 1 && thr >= 0, "hamx_loop_score_dev: bad sizes (nframes <= 65535 per call)");
    ORBX_REQUIRE(d_scores != nullptr || nframes == 0, "hamx_loop_score_dev: NULL output");
    ORBX_CUDA(cudaSetDevice(h->device));
    if (nframes) ORBX_CUDA(cudaMemsetAsync(d_scores, 0, (size_t)nframes * sizeof(int32_t), h->stream));
    if (nframes && nq) {
        ORBX_REQUIRE(d_q && d_frames && d_counts, "hamx_loop_score_dev: NULL pointer");
        if ((((uintptr_t)d_q) | ((uintptr_t)d_frames)) & 15) { set_error("hamx_loop_score_dev: descriptor pointers must be 16-byte aligned"); return ORBX_E_ALIGN; }
        const bool tc = h->kernel_mode == HAMX_KERNEL_TENSOR ||
                        (h->kernel_mode == HAMX_KERNEL_AUTO && (int64_t)nq * cap * nframes >= (1ll << 24) && thr <= 256);
        if (tc) {
            // tensor-core formulation (hamming_tc.cuh): the stored frames are expanded into operand images, at most ~256 MB at a time
            if (!h->loop_tc_ready) {
                ORBX_CUDA(cudaFuncSetAttribute(k_loop_score_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
                h->loop_tc_ready = true;
            }
            const int tpf = (cap + TC_TN - 1) / TC_TN;
            const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nframes, ((size_t)256 << 20) / ((size_t)tpf * TC_TILE_BYTES)));
            int rc = grow(&h->d_texp, &h->texp_bytes, (size_t)chunk * tpf * TC_TILE_BYTES);
            if (rc) return rc;
            const unsigned int nqb = (unsigned int)((nq + TC_QB - 1) / TC_QB);
            for (int c0 = 0; c0 < nframes; c0 += chunk) {
                const int nf = std::min(chunk, nframes - c0);
                const long long per_frame = (long long)tpf * (TC_TILE_BYTES / 16);
                k_expand_frames<<<dim3((unsigned int)((per_frame + 255) / 256), (unsigned int)nf), 256, 0, h->stream>>>(
                    d_frames + (size_t)c0 * cap * 32, d_counts + c0, cap, tpf, reinterpret_cast<uint4*>(h->d_texp));
                ORBX_CUDA(cudaGetLastError());
                // two CTAs' worth of work per SM, each walking a run of frames with its queries resident
                int fpc = (int)std::max<int64_t>(1, ((int64_t)nf * nqb + 2 * h->sm_count - 1) / (2 * h->sm_count));
                if (fpc > nf) fpc = nf;
                const dim3 grid(nqb, (unsigned int)((nf + fpc - 1) / fpc));
                k_loop_score_tc<<<grid, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(d_q, nq, h->d_texp, d_counts + c0, cap, nf, tpf, fpc, n, thr,
                                                                                d_scores + c0);
                ORBX_CUDA(cudaGetLastError());
            }
        } else {
            const dim3 grid((unsigned)((nq + HT_QB - 1) / HT_QB), (unsigned)nframes);
            k_loop_score<<<grid, HT_THREADS, 0, h->stream>>>(d_q, nq, d_frames, d_counts, cap, n, (uint32_t)thr, d_scores);
            ORBX_CUDA(cudaGetLastError());
        }
    }
    if (d_best) return hamx_loop_best_dev(h, d_scores, nframes, d_best);
    return ORBX_OK;
}

extern "C" int hamx_loop_best_dev(hamx_handle h, const int32_t* d_scores, int nframes, int32_t* d_best)
{
    ORBX_REQUIRE(h != nullptr && d_best != nullptr && nframes >= 0 && (nframes == 0 || d_scores), "hamx_loop_best_dev: bad arguments");
    ORBX_CUDA(cudaSetDevice(h->device));
    k_loop_best<<<1, 256, 0, h->stream>>>(d_scores, nframes, d_best);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int hamx_nbest_dev(hamx_handle h, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int n, int32_t* d_dist, int32_t* d_idx)
{
    ORBX_REQUIRE(h != nullptr, "hamx_nbest_dev: NULL handle");
    ORBX_REQUIRE(nq >= 0 && nt >= 0 && n >= 1 && n <= 16, "hamx_nbest_dev: bad sizes (1 <= n <= 16)");
    if (nq == 0) return ORBX_OK;
    ORBX_REQUIRE(d_q && d_dist && d_idx && (nt == 0 || d_t), "hamx_nbest_dev: NULL pointer");
    if ((((uintptr_t)d_q) | ((uintptr_t)d_t)) & 15) { set_error("hamx_nbest_dev: descriptor pointers must be 16-byte aligned"); return ORBX_E_ALIGN; }
    ORBX_CUDA(cudaSetDevice(h->device));
    const unsigned int blocks = (unsigned int)((nq + HT_THREADS - 1) / HT_THREADS);
    if (n <= 4) k_nbest<4><<<blocks, HT_THREADS, 0, h->stream>>>(d_q, nq, d_t, nt, n, d_dist, d_idx);
    else if (n <= 10) k_nbest<10><<<blocks, HT_THREADS, 0, h->stream>>>(d_q, nq, d_t, nt, n, d_dist, d_idx);
    else k_nbest<16><<<blocks, HT_THREADS, 0, h->stream>>>(d_q, nq, d_t, nt, n, d_dist, d_idx);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

// Host-buffer form of LoopCloser::DetectLoop's scoring: descriptors of the current frame against nframes stored frames.
extern "C" int hamx_loop_score(hamx_handle h, const uint8_t* q, int nq, const uint8_t* frames, const int32_t* counts, int nframes, int cap, int n,
                               int thr, int32_t* scores, int32_t* best_frame)
{
    ORBX_REQUIRE(h != nullptr && best_frame != nullptr, "hamx_loop_score: NULL argument");
    ORBX_REQUIRE(nq >= 0 && nframes >= 0 && cap >= 1, "hamx_loop_score: bad sizes");
    *best_frame = -1;
    if (nframes == 0) return ORBX_OK;
    ORBX_REQUIRE(scores && counts && frames && (nq == 0 || q), "hamx_loop_score: NULL pointer");
    ORBX_CUDA(cudaSetDevice(h->device));
    int rc = grow(&h->d_q, &h->q_bytes, (size_t)nq * 32 + 32);
    if (!rc) rc = grow(&h->d_t, &h->t_bytes, (size_t)nframes * cap * 32 + 32);
    if (!rc) rc = grow(&h->d_counts, &h->counts_bytes, ((size_t)2 * nframes + 2) * sizeof(int32_t));
    if (rc) return rc;
    int32_t* d_cnt = h->d_counts;
    int32_t* d_sc = h->d_counts + nframes;
    if (nq) ORBX_CUDA(cudaMemcpyAsync(h->d_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(h->d_t, frames, (size_t)nframes * cap * 32, cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(d_cnt, counts, (size_t)nframes * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    int32_t* d_best = reinterpret_cast<int32_t*>(h->d_ngood);       // 8 bytes: {frame, score}
    rc = hamx_loop_score_dev(h, h->d_q, nq, h->d_t, d_cnt, nframes, cap, n, thr, d_sc, d_best);
    if (rc) return rc;
    int32_t best[2] = { -1, 0 };
    ORBX_CUDA(cudaMemcpyAsync(scores, d_sc, (size_t)nframes * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(best, d_best, sizeof(best), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    *best_frame = best[0];
    return ORBX_OK;
}

// NBestMatches(descriptors1, descriptors2, n, distances, indices) on host buffers.
extern "C" int hamx_nbest(hamx_handle h, const uint8_t* q, int nq, const uint8_t* t, int nt, int n, int32_t* dist, int32_t* idx)
{
    ORBX_REQUIRE(h != nullptr, "hamx_nbest: NULL handle");
    ORBX_REQUIRE(nq >= 0 && nt >= 0 && n >= 1 && n <= 16, "hamx_nbest: bad sizes (1 <= n <= 16)");
    if (nq == 0) return ORBX_OK;
    ORBX_REQUIRE(q && dist && idx && (nt == 0 || t), "hamx_nbest: NULL pointer");
    ORBX_CUDA(cudaSetDevice(h->device));
    int rc = upload_sets(h, q, nq, t, nt);
    if (!rc) rc = grow(&h->d_dm, &h->dm_bytes, (size_t)nq * n * 2 * sizeof(int32_t));
    if (rc) return rc;
    int32_t* d_dist = reinterpret_cast<int32_t*>(h->d_dm);
    int32_t* d_idx = d_dist + (size_t)nq * n;
    rc = hamx_nbest_dev(h, h->d_q, nq, h->d_t, nt, n, d_dist, d_idx);
    if (rc) return rc;
    ORBX_CUDA(cudaMemcpyAsync(dist, d_dist, (size_t)nq * n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(idx, d_idx, (size_t)nq * n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

// ------------------------------------------------------------------------------------------------ peer-memory path
static size_t p2p_flag_bytes(int world) { return align_up((size_t)2 * world * sizeof(unsigned int), 256); }

static void p2p_set_view(hamx_handle h, int r, void* base)
{
    h->pv.flags[r] = reinterpret_cast<unsigned int*>(base);
    h->pv.gather[r] = reinterpret_cast<hamx_top2*>(reinterpret_cast<uint8_t*>(base) + p2p_flag_bytes(h->pv.world));
}

extern "C" int hamx_p2p_export(hamx_handle h, int64_t nq_max, int world, int rank, uint8_t* ipc_handle, void** local_base)
{
    ORBX_REQUIRE(h != nullptr, "hamx_p2p_export: NULL handle");
    ORBX_REQUIRE(world >= 1 && world <= HT_MAX_WORLD && rank >= 0 && rank < world && nq_max >= 1 && nq_max < (1ll << 31),
                 "hamx_p2p_export: bad world %d / rank %d / nq_max %lld", world, rank, (long long)nq_max);
    ORBX_REQUIRE(h->p2p_buf == nullptr, "hamx_p2p_export: already exported; call hamx_p2p_close first");
    ORBX_CUDA(cudaSetDevice(h->device));
    memset(&h->pv, 0, sizeof(h->pv));
    h->pv.world = world; h->pv.rank = rank; h->pv.nq_max = nq_max;
    h->p2p_bytes = p2p_flag_bytes(world) + (size_t)2 * world * (size_t)nq_max * sizeof(hamx_top2);
    cudaError_t e = cudaMalloc(&h->p2p_buf, h->p2p_bytes);
    if (e != cudaSuccess) { h->p2p_buf = nullptr; set_error("hamx_p2p_export: cudaMalloc(%zu) failed: %s", h->p2p_bytes, cudaGetErrorString(e)); return ORBX_E_ALLOC; }
    ORBX_CUDA_OR(cudaMemset(h->p2p_buf, 0, p2p_flag_bytes(world)), hamx_p2p_close(h));     // epoch 0 = nothing published
    ORBX_CUDA_OR(cudaMalloc((void**)&h->pv.done, 256), hamx_p2p_close(h));
    ORBX_CUDA_OR(cudaMemset(h->pv.done, 0, 256), hamx_p2p_close(h));
    ORBX_CUDA_OR(cudaDeviceSynchronize(), hamx_p2p_close(h));
    p2p_set_view(h, rank, h->p2p_buf);
    if (ipc_handle) {
        cudaIpcMemHandle_t ih;
        ORBX_CUDA(cudaIpcGetMemHandle(&ih, h->p2p_buf));
        static_assert(sizeof(ih) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(ipc_handle, &ih, sizeof(ih));
    }
    if (local_base) *local_base = h->p2p_buf;
    return ORBX_OK;
}

extern "C" int hamx_p2p_import(hamx_handle h, const uint8_t* ipc_handles)
{
    ORBX_REQUIRE(h != nullptr && h->p2p_buf != nullptr && ipc_handles != nullptr, "hamx_p2p_import: export first");
    ORBX_CUDA(cudaSetDevice(h->device));
    for (int r = 0; r < h->pv.world; r++) {
        if (r == h->pv.rank) continue;
        cudaIpcMemHandle_t ih;
        memcpy(&ih, ipc_handles + (size_t)r * sizeof(ih), sizeof(ih));
        void* base = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&base, ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { set_error("hamx_p2p_import: cannot map rank %d's buffer: %s", r, cudaGetErrorString(e)); return ORBX_E_CUDA; }
        h->p2p_opened[r] = base;
        p2p_set_view(h, r, base);
    }
    return ORBX_OK;
}

extern "C" int hamx_p2p_import_ptrs(hamx_handle h, void* const* peer_bases)
{
    ORBX_REQUIRE(h != nullptr && h->p2p_buf != nullptr && peer_bases != nullptr, "hamx_p2p_import_ptrs: export first");
    for (int r = 0; r < h->pv.world; r++) {
        if (r == h->pv.rank) continue;
        ORBX_REQUIRE(peer_bases[r] != nullptr, "hamx_p2p_import_ptrs: NULL base for rank %d", r);
        p2p_set_view(h, r, peer_bases[r]);
    }
    return ORBX_OK;
}

extern "C" int hamx_p2p_close(hamx_handle h)
{
    if (!h) return ORBX_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int r = 0; r < HT_MAX_WORLD; r++)
        if (h->p2p_opened[r]) { cudaIpcCloseMemHandle(h->p2p_opened[r]); h->p2p_opened[r] = nullptr; }
    if (h->pv.done) cudaFree(h->pv.done);
    if (h->p2p_buf) cudaFree(h->p2p_buf);
    h->p2p_buf = nullptr;
    memset(&h->pv, 0, sizeof(h->pv));
    return ORBX_OK;
}

static int p2p_ready(hamx_handle h, int64_t nq, const char* fn)
{
    ORBX_REQUIRE(h != nullptr, "%s: NULL handle", fn);
    ORBX_REQUIRE(h->p2p_buf != nullptr, "%s: call hamx_p2p_export / hamx_p2p_import first", fn);
    for (int r = 0; r < h->pv.world; r++) ORBX_REQUIRE(h->pv.gather[r] != nullptr, "%s: rank %d's buffer has not been imported", fn, r);
    ORBX_REQUIRE(nq >= 1 && nq <= h->pv.nq_max, "%s: nq %lld outside [1, nq_max=%lld]", fn, (long long)nq, (long long)h->pv.nq_max);
    return ORBX_OK;
}

// Collective, phase 1: local top-2 over this rank's train shard, scattered into every rank's gather buffer by the
// matching kernel itself.  All ranks must call it with the same nq, the same number of times.
extern "C" int hamx_knn2_p2p_scatter_dev(hamx_handle h, const uint8_t* d_q, int64_t nq, const uint8_t* d_t, int64_t nt, int64_t train_offset)
{
    int rc = p2p_ready(h, nq, "hamx_knn2_p2p_scatter_dev");
    if (rc) return rc;
    ORBX_REQUIRE(nt >= 0 && train_offset >= 0 && nt + train_offset < (1ll << 31), "hamx_knn2_p2p_scatter_dev: indices must fit int32");
    ORBX_REQUIRE(d_q && (nt == 0 || d_t), "hamx_knn2_p2p_scatter_dev: NULL pointer");
    if ((((uintptr_t)d_q) | ((uintptr_t)d_t)) & 15) { set_error("hamx_knn2_p2p_scatter_dev: device pointers must be 16-byte aligned"); return ORBX_E_ALIGN; }
    ORBX_CUDA(cudaSetDevice(h->device));
    h->pv.epoch += 1;
    if (h->pv.epoch == 0) h->pv.epoch = 1;
    if (nt > 0 && nt <= (1ll << HT_IDX_BITS)) return launch_chunk(h, d_q, nq, d_t, (int)nt, train_offset, nullptr, &h->pv);
    // empty or multi-launch shard: compute locally, then scatter
    rc = grow(&h->d_p2p_local, &h->p2p_local_bytes, (size_t)nq * sizeof(hamx_top2));
    if (rc) return rc;
    rc = hamx_knn2_dev(h, d_q, nq, d_t, nt, train_offset, h->d_p2p_local);
    if (rc) return rc;
    const int blocks = (int)std::min<int64_t>((nq + 255) / 256, (int64_t)h->sm_count * 8);
    k_scatter_p2p<<<blocks, 256, 0, h->stream>>>(h->d_p2p_local, nq, h->pv);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

// Collective, phase 2: wait for every rank's scatter of the current call, merge -> d_out[nq] (identical on all ranks).
extern "C" int hamx_p2p_merge_dev(hamx_handle h, int64_t nq, hamx_top2* d_out)
{
    int rc = p2p_ready(h, nq, "hamx_p2p_merge_dev");
    if (rc) return rc;
    ORBX_REQUIRE(d_out != nullptr && h->pv.epoch != 0, "hamx_p2p_merge_dev: no scatter in flight or NULL output");
    ORBX_CUDA(cudaSetDevice(h->device));
    k_merge_top2_p2p<<<(unsigned int)((nq + 255) / 256), 256, 0, h->stream>>>(nq, d_out, h->pv);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int hamx_knn2_p2p_dev(hamx_handle h, const uint8_t* d_q, int64_t nq, const uint8_t* d_t, int64_t nt, int64_t train_offset,
                                 hamx_top2* d_out)
{
    ORBX_REQUIRE(h != nullptr && d_out != nullptr, "hamx_knn2_p2p_dev: NULL argument");
    // small query sets on a shard that one launch covers: the matching kernel also waits for the world and merges
    const bool fuse = nq <= 16384 && nt > 0 && nt <= (1ll << HT_IDX_BITS) && h->p2p_buf != nullptr;
    if (fuse) { h->pv.merge_out = d_out; h->pv.merge_nq = nq; }
    int rc = hamx_knn2_p2p_scatter_dev(h, d_q, nq, d_t, nt, train_offset);
    h->pv.merge_out = nullptr;
    h->pv.merge_nq = 0;
    if (rc || fuse) return rc;
    return hamx_p2p_merge_dev(h, nq, d_out);
}

extern "C" int hamx_popc_peak(int device, double* gpopc_per_s, double* elapsed_ms)
{
    ORBX_REQUIRE(gpopc_per_s != nullptr, "hamx_popc_peak: NULL output");
    ORBX_CUDA(cudaSetDevice(device));
    int sms = 0;
    ORBX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    uint32_t* sink = nullptr;
    ORBX_CUDA(cudaMalloc((void**)&sink, 256));
    cudaEvent_t e0, e1;
    ORBX_CUDA(cudaEventCreate(&e0));
    ORBX_CUDA(cudaEventCreate(&e1));
    const int iters = 1 << 16, blocks = sms * 2, threads = 1024;
    k_popc_peak<<<blocks, threads>>>(sink, iters / 16, 1u);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        ORBX_CUDA(cudaEventRecord(e0));
        k_popc_peak<<<blocks, threads>>>(sink, iters, 12345u + rep);
        ORBX_CUDA(cudaEventRecord(e1));
        ORBX_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        ORBX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    ORBX_CUDA(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    double popc = (double)blocks * threads * (double)iters * 8.0;
    *gpopc_per_s = popc / (best * 1e-3) / 1e9;
    if (elapsed_ms) *elapsed_ms = best;
    return ORBX_OK;
}
