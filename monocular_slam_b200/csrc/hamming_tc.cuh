// hamming_tc.cuh -- K6 on the 5th-generation tensor cores (included by hamming.cu; sm_100a only).
//
// The 256-bit Hamming distance is a contraction: with the bits mapped to +-1, s = sum_k q'_k t'_k = 256 - 2 ham(q, t), exact in
// int32.  So the distance matrix of a block of queries against a tile of train rows is one int8 GEMM with K = 256, and
// tcgen05.mma (kind::i8, operands in shared memory, accumulators in tensor memory) produces 128 x 128 distances per 8
// instructions while the integer pipes only have to fold them into the per-query top-2 (2.5 min/max per distance instead of
// the ~20 operations the XOR/POPC formulation spends on computing each one).  Results are identical to k_hamming_knn2: the
// train side carries -64 t'_k, and one more K step (query side 1, train side the row's index n inside its tile, 0..127)
// makes the accumulator itself the key  -64 s + n = 128 ham - 16384 + n,  which orders like (distance, trainIdx) -- the
// epilogue does not touch the value before comparing it.
//
// Layout.  Operands are K-major, no swizzle ("interleaved" canonical layout): 8 rows x 16 bytes form a contiguous 128-byte core
// matrix; a 128-row x 288-byte tile (256 bit bytes + the 32-byte index step) is stored as [row / 8][k / 16][row % 8][16 B] =
// 36 KB, i.e. leading-dimension (K) byte offset 128, stride-dimension (8-row group) byte offset 2304; K step j of an MMA
// (32 bytes) starts 256 j bytes in.  k_expand_train writes the train set in exactly this image, one contiguous 36 KB block per
// tile, so a tile arrives with ONE bulk copy (cp.async.bulk + mbarrier, no tensor map).  The CTA's 256 queries are expanded
// into shared memory by the CTA itself.
//
// CTA = 18 warps: warps 0-15 fold accumulators (warp w owns TMEM lanes 32 (w % 4) .. +31 of query block (w / 4) % 2, columns
// [64 (w / 8), +64) of every tile), warp 16 lane 0 streams train tiles (3-stage ring), warp 17 lane 0 issues the MMAs and owns
// the TMEM allocation (512 columns: 2 query blocks x 128 columns x 2 accumulator stages, so the MMAs of tile i+1 run while
// tile i is being folded).  Tensor memory can be read at 64 bytes per clock and SM, i.e. 16 accumulators: that, not the MMA
// rate and not the integer pipes, is what bounds this formulation with 32-bit accumulators.
#pragma once

namespace orbx {
namespace {

constexpr int TC_EPI_WARPS = 16;
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;
constexpr int TC_QB = 256;                 // queries per CTA (two MMA row blocks of 128)
static_assert(TC_QB == HT_QB, "the peer-memory epilogue (p2p_publish) merges one block of HT_QB queries per writer");
constexpr int TC_TN = 128;                 // train rows per tile (MMA N)
constexpr int TC_KC = 18;                  // 16-byte K chunks per row: 16 of descriptor bits, 1 with the index byte, 1 of zeros
constexpr int TC_KSTEPS = TC_KC / 2;       // MMA K steps (32 bytes each)
constexpr uint32_t TC_LBO = 128, TC_SBO = TC_KC * 128;
constexpr int TC_TILE_BYTES = 16 * (int)TC_SBO;   // 36 KB
#ifndef TC_STAGES_N
#define TC_STAGES_N 3
#endif
constexpr int TC_STAGES = TC_STAGES_N;
constexpr int TC_NONE_KEY = 0x7FFFFFFF;
// kind::i8 instruction descriptor: D = s32 (bits 4-5 = 2), A and B signed 8-bit (bits 7-9, 10-12 = 1), both K-major (bits 15, 16 = 0),
// N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t TC_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_TN >> 3) << 17) | ((128u >> 4) << 24);
constexpr size_t TC_SMEM_BYTES = (size_t)2 * TC_TILE_BYTES + (size_t)TC_STAGES * TC_TILE_BYTES + 1024;
static_assert(TC_SMEM_BYTES <= 227 * 1024, "operand images must fit the shared memory of one SM");

// 16 descriptor bits -> 16 bytes (bit i of the pair of bytes -> byte i): queries +1 / -1, train rows -64 / +64
template <bool TRAIN>
__device__ __forceinline__ uint32_t expand_nibble(uint32_t nib)
{
    const uint32_t w = (nib * 0x00204081u) & 0x01010101u;          // byte i = bit i of the nibble
    if (TRAIN) return 0x40404040u + w * 0x80u;                      // 1 -> 0xC0 (-64), 0 -> 0x40 (+64); no carries between bytes
    return 0x01010101u | ((w ^ 0x01010101u) * 0xFEu);               // 1 -> 0x01 (+1), 0 -> 0xFF (-1)
}
template <bool TRAIN>
__device__ __forceinline__ uint4 expand_bits16(uint32_t bits)
{
    return make_uint4(expand_nibble<TRAIN>(bits & 15u), expand_nibble<TRAIN>((bits >> 4) & 15u), expand_nibble<TRAIN>((bits >> 8) & 15u),
                      expand_nibble<TRAIN>((bits >> 12) & 15u));
}

// out: ceil(nt / 128) tiles of 36 KB in the shared-memory image described above; rows beyond nt are zero (masked by the consumer).
// With a pair table (batched frame pairs) blockIdx.y selects the pair, whose image starts pair_stride 16-byte units further on.
__global__ void __launch_bounds__(256) k_expand_train(const uint8_t* __restrict__ t, int nt, const hamx_pair* __restrict__ pairs,
                                                      size_t pair_stride, uint4* __restrict__ out)
{
    if (pairs) {
        const hamx_pair pd = pairs[blockIdx.y];
        t = pd.nq > 0 ? pd.t : nullptr;
        nt = pd.nq > 0 ? pd.nt : 0;
        out += (size_t)blockIdx.y * pair_stride;
    }
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte unit each
    const long long ntiles = (nt + TC_TN - 1) / TC_TN;
    if (u >= ntiles * (TC_TILE_BYTES / 16)) return;
    const int tile = (int)(u / (TC_TILE_BYTES / 16)), r = (int)(u % (TC_TILE_BYTES / 16));
    const int n1 = r / (TC_KC * 8), kc = (r / 8) % TC_KC, n0 = r & 7;
    const int nl = n1 * 8 + n0, row = tile * TC_TN + nl;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < nt) {
        if (kc < 16) v = expand_bits16<true>(*reinterpret_cast<const uint16_t*>(t + (size_t)row * 32 + 2 * kc));
        else if (kc == 16) v.x = (uint32_t)nl;       // byte 0 of the index step: the row's index inside the tile (query side: 1)
    }
    out[u] = v;
}

__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr)
{
    // start address >> 4 | LBO >> 4 at bit 16 | SBO >> 4 at bit 32 | descriptor version 1 at bit 46 | no swizzle (bits 61-63 = 0)
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(TC_LBO >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32(uint32_t taddr, int (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 64 consecutive columns of this thread's TMEM lane as 32 registers: .pack::16b puts the low halves of two adjacent 32-bit
// columns into one register (column 2r in bits 0-15, column 2r+1 in bits 16-31).  The keys fit 16 signed bits
// (-16384 .. 16511), so nothing is lost and the min/max below handle two columns per instruction.
__device__ __forceinline__ void tc_ld64_pack16(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

#ifndef TC_PACK16
#define TC_PACK16 1
#endif

// Same launch geometry and split / merge protocol as k_hamming_knn2 (grid = (query blocks of 256, train splits, pairs)); `texp`
// is the expanded train set of k_expand_train.  With `pairs` the launch handles pair blockIdx.z of a device-resident table.
template <bool P2P>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_hamming_tc(const uint8_t* __restrict__ q_, int64_t nq, const uint8_t* __restrict__ texp, int nt, const hamx_pair* __restrict__ pairs,
             size_t texp_pair_stride, int tiles_per_split, uint2* partial, size_t nq_stride, unsigned int* arrivals, int qblocks_stride,
             hamx_top2* __restrict__ out, size_t out_stride, int64_t idx_offset, const __grid_constant__ P2PView pv)
{
    if (pairs) {
        const hamx_pair pd = pairs[blockIdx.z];
        q_ = pd.q;
        nq = pd.nq;
        nt = pd.nt;
        if ((int64_t)blockIdx.x * TC_QB >= nq) return;    // batched launches are sized for the largest pair
        texp += (size_t)blockIdx.z * texp_pair_stride;
        out += (size_t)blockIdx.z * out_stride;
        partial += (size_t)blockIdx.z * gridDim.y * nq_stride;
        arrivals += (size_t)blockIdx.z * qblocks_stride;
    }
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t s_full[TC_STAGES], s_empty[TC_STAGES], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_last;
    __shared__ uint2 s_half[TC_QB];       // top-2 of the warps that fold columns 64..127, handed to their partners

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)tc_smem + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                  // 2 x 36 KB: the CTA's queries, expanded
    uint8_t* sB = smem + 2 * TC_TILE_BYTES;              // TC_STAGES x 36 KB: train tiles

    const int ntiles_all = (nt + TC_TN - 1) / TC_TN;
    const int tile_begin = blockIdx.y * tiles_per_split;
    const int ntiles = max(min(tile_begin + tiles_per_split, ntiles_all) - tile_begin, 0);
    const int nsplit = gridDim.y;
    const int64_t qbase = (int64_t)blockIdx.x * TC_QB;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&s_tfull[a], 1); mbar_init(&s_tempty[a], TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == TC_EPI_WARPS + 1) {      // the allocating warp also frees; all 512 columns (the kernel runs one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the CTA's queries -> +-1 bytes in the operand image, index step = (1, 0, ..); rows beyond nq: zeros, never written out
    {
        const uint8_t* __restrict__ q = q_;
        for (int u = tid; u < TC_QB * TC_KC; u += TC_THREADS) {
            const int row = u / TC_KC, kc = u - row * TC_KC;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (qbase + row < nq) {
                if (kc < 16) v = expand_bits16<false>(*reinterpret_cast<const uint16_t*>(q + (size_t)(qbase + row) * 32 + 2 * kc));
                else if (kc == 16) v.x = 1u;
            }
            const int blk = row >> 7, r = row & 127;
            *reinterpret_cast<uint4*>(sA + blk * TC_TILE_BYTES + (r >> 3) * TC_SBO + kc * TC_LBO + (r & 7) * 16) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core's reads
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    uint32_t gk0 = HT_NONE, gk1 = HT_NONE;       // the folding warps' running top-2 in k_hamming_knn2's key format

    if (wid == TC_EPI_WARPS) {
        if (lane == 0) {
            for (int i = 0; i < ntiles; i++) {
                const int s = i % TC_STAGES;
                mbar_wait(&s_empty[s], ((uint32_t)(i / TC_STAGES) & 1u) ^ 1u);
                mbar_expect_tx(&s_full[s], (uint32_t)TC_TILE_BYTES);
                tma_bulk_g2s(sB + s * TC_TILE_BYTES, texp + (size_t)(tile_begin + i) * TC_TILE_BYTES, (uint32_t)TC_TILE_BYTES, &s_full[s]);
            }
        }
    } else if (wid == TC_EPI_WARPS + 1) {
        if (lane == 0) {
            const uint64_t adesc0 = tc_smem_desc(smem_u32(sA)), adesc1 = tc_smem_desc(smem_u32(sA + TC_TILE_BYTES));
            for (int i = 0; i < ntiles; i++) {
                const int s = i % TC_STAGES, a = i & 1;
                mbar_wait(&s_full[s], (uint32_t)(i / TC_STAGES) & 1u);
                mbar_wait(&s_tempty[a], ((uint32_t)(i >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint64_t bdesc = tc_smem_desc(smem_u32(sB + s * TC_TILE_BYTES));
                const uint32_t d0 = tmem_base + (uint32_t)(a * 256), d1 = d0 + 128u;
#pragma unroll
                for (int k = 0; k < TC_KSTEPS; k++) {     // K step of 32 bytes = 2 core matrices = 256 bytes further on (16 in descriptor units)
                    tc_mma_i8(d0, adesc0 + (uint64_t)(16 * k), bdesc + (uint64_t)(16 * k), TC_IDESC, k > 0);
                    tc_mma_i8(d1, adesc1 + (uint64_t)(16 * k), bdesc + (uint64_t)(16 * k), TC_IDESC, k > 0);
                }
                tc_commit(&s_empty[s]);           // the stage may be refilled once these MMAs have read it
                tc_commit(&s_tfull[a]);           // and the accumulators are complete
            }
        }
    } else {
        const int blk = (wid >> 2) & 1, half = wid >> 3;
        const uint32_t lane_base = (uint32_t)(32 * (wid & 3)) << 16;
        for (int i = 0; i < ntiles; i++) {
            const int a = i & 1;
            mbar_wait(&s_tfull[a], (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const int tile = tile_begin + i;
            const int valid = min(TC_TN, nt - tile * TC_TN);
            int t0 = TC_NONE_KEY, t1 = TC_NONE_KEY;
            if (TC_PACK16 && half * 64 + 64 <= valid) {
                // the warp's 64 columns as 32 registers of two 16-bit keys: each half-word lane keeps its own top-2 (even / odd
                // columns) in two independent chains -- 5 packed min/max per 8 keys
                uint32_t p[32];
                tc_ld64_pack16(tmem_base + lane_base + (uint32_t)(a * 256 + blk * 128 + half * 64), p);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_tempty[a]);      // the registers hold everything this warp needs from the stage
                uint32_t a0 = 0x7FFF7FFFu, a1 = 0x7FFF7FFFu, b0 = 0x7FFF7FFFu, b1 = 0x7FFF7FFFu;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const uint32_t lo = __vmins2(p[j], p[j + 1]), hi = __vmaxs2(p[j], p[j + 1]);
                    const uint32_t m = __vmaxs2(a0, lo);
                    a0 = __vmins2(a0, lo);
                    a1 = __vimin3_s16x2(a1, m, hi);
                    const uint32_t lo2 = __vmins2(p[j + 2], p[j + 3]), hi2 = __vmaxs2(p[j + 2], p[j + 3]);
                    const uint32_t m2 = __vmaxs2(b0, lo2);
                    b0 = __vmins2(b0, lo2);
                    b1 = __vimin3_s16x2(b1, m2, hi2);
                }
                const uint32_t m = __vmaxs2(a0, b0);
                a0 = __vmins2(a0, b0);
                a1 = __vimin3_s16x2(a1, b1, m);
                // (a0 <= a1) per half-word lane -> the two smallest of the four
                const int el = (int)(short)(a0 & 0xFFFFu), eh = (int)a0 >> 16, fl = (int)(short)(a1 & 0xFFFFu), fh = (int)a1 >> 16;
                t0 = min(el, eh);
                t1 = __vimin3_s32(fl, fh, max(el, eh));
                if (t1 == 0x7FFF) t1 = TC_NONE_KEY;
                if (t0 == 0x7FFF) t0 = TC_NONE_KEY;
            } else {
            // two independent insertion chains (even / odd column pairs) keep the min/max pipe busy; merged per tile
            int u0 = TC_NONE_KEY, u1 = TC_NONE_KEY;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int col0 = half * 64 + c * 32;
                int v[32];
                tc_ld32(tmem_base + lane_base + (uint32_t)(a * 256 + blk * 128 + col0), v);
                if (col0 + 32 <= valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const int lo = min(v[j], v[j + 1]), hi = max(v[j], v[j + 1]);
                        const int m = max(t0, lo);
                        t0 = min(t0, lo);
                        t1 = __vimin3_s32(t1, m, hi);
                        const int lo2 = min(v[j + 2], v[j + 3]), hi2 = max(v[j + 2], v[j + 3]);
                        const int m2 = max(u0, lo2);
                        u0 = min(u0, lo2);
                        u1 = __vimin3_s32(u1, m2, hi2);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int x = col0 + j < valid ? v[j] : TC_NONE_KEY;
                        const int m = max(t0, x);
                        t0 = min(t0, x);
                        t1 = min(t1, m);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_tempty[a]);
            {   // (t0 <= t1) and (u0 <= u1) -> the two smallest of the four
                const int lo = min(t0, u0), m = max(t0, u0);
                t1 = __vimin3_s32(t1, u1, m);
                t0 = lo;
            }
            }
            // tile-local keys 128 ham - 16384 + n  ->  ham << 23 | (tile * 128 + n), folded into the running top-2
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int k = e ? t1 : t0;
                if (k != TC_NONE_KEY) {
                    const uint32_t u = (uint32_t)(k + 16384);
                    top2_insert(gk0, gk1, ((u >> 7) << HT_IDX_BITS) + ((uint32_t)tile * TC_TN + (u & 127u)));
                }
            }
        }
        if (half) s_half[tid - 256] = make_uint2(gk0, gk1);
    }
    tc_fence_before();
    __syncthreads();
    if (wid == TC_EPI_WARPS + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }

    // ---- results: thread -> query (threads of warps 0-7 hold one; their partners' halves come through shared memory)
    const bool holder = wid < 8;
    const int64_t qi = qbase + tid;            // warp w < 8: block w / 4, lanes 32 (w % 4) ..  ==  row tid of the CTA's 256
    if (holder) {
        const uint2 o = s_half[tid];
        top2_insert(gk0, gk1, o.x);
        top2_insert(gk0, gk1, o.y);
    }
    if (nsplit == 1) {
        if (holder && qi < nq) {
            const hamx_top2 v = decode_top2(gk0, gk1, idx_offset);
            if (P2P) p2p_store(pv, qi, v);
            else out[qi] = v;
        }
        if (P2P) p2p_publish(pv, gridDim.x);
        return;
    }
    if (holder && qi < nq) __stcg(&partial[(size_t)blockIdx.y * nq_stride + qi], make_uint2(gk0, gk1));
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int prev = atomicAdd(&arrivals[blockIdx.x], 1u);
        s_last = (prev == (unsigned int)(nsplit - 1));
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (holder && qi < nq) {
        uint32_t m0 = HT_NONE, m1 = HT_NONE;
        int s = 0;
        for (; s + 8 <= nsplit; s += 8) {      // independent loads in flight, as in k_hamming_knn2
            uint2 p[8];
#pragma unroll
            for (int u = 0; u < 8; u++) p[u] = __ldcg(&partial[(size_t)(s + u) * nq_stride + qi]);
#pragma unroll
            for (int u = 0; u < 8; u++) { top2_insert(m0, m1, p[u].x); top2_insert(m0, m1, p[u].y); }
        }
        for (; s < nsplit; s++) {
            const uint2 p = __ldcg(&partial[(size_t)s * nq_stride + qi]);
            top2_insert(m0, m1, p.x);
            top2_insert(m0, m1, p.y);
        }
        const hamx_top2 v = decode_top2(m0, m1, idx_offset);
        if (P2P) p2p_store(pv, qi, v);
        else out[qi] = v;
    }
    if (tid == 0) arrivals[blockIdx.x] = 0;   // ready for the next launch
    if (P2P) p2p_publish(pv, gridDim.x);      // one merging CTA per query block
}

// ---- loop-closure candidate scoring on the same machinery (k_loop_score's contract, hamming.cu) ---------------------------------
// grid = (query blocks of 256, frame chunks).  A CTA keeps its 256 expanded queries in shared memory and walks the stored frames
// [f0, f0 + frames_per_cta) of its chunk as ONE flat stream of train tiles (every frame occupies tiles_per_frame tiles of the
// expanded image, so the producer and the MMA issuer do not see frame borders).  The folding warps count, per query, the keys
// below the threshold (key < 128 thr - 16384  <=>  distance < thr, because the index part is < 128); at the last tile of a
// frame the two warps that share a query add their counts in shared memory, the 16 warps meet at a named barrier, and the
// count is clipped at n, reduced and added to the frame's score.
__global__ void __launch_bounds__(TC_THREADS, 1)
k_loop_score_tc(const uint8_t* __restrict__ q_, int nq, const uint8_t* __restrict__ texp, const int32_t* __restrict__ counts, int cap, int nframes,
                int tiles_per_frame, int frames_per_cta, int nbest, int thr, int32_t* __restrict__ scores)
{
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t s_full[TC_STAGES], s_empty[TC_STAGES], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_cnt[2][TC_QB];

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)tc_smem + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + 2 * TC_TILE_BYTES;
    const int f0 = blockIdx.y * frames_per_cta;
    const int nf = max(min(f0 + frames_per_cta, nframes) - f0, 0);
    const int ntiles = nf * tiles_per_frame;
    const int qbase = blockIdx.x * TC_QB;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&s_tfull[a], 1); mbar_init(&s_tempty[a], TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == TC_EPI_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int u = tid; u < 2 * TC_QB; u += TC_THREADS) (&s_cnt[0][0])[u] = 0;
    for (int u = tid; u < TC_QB * TC_KC; u += TC_THREADS) {
        const int row = u / TC_KC, kc = u - row * TC_KC;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (qbase + row < nq) {
            if (kc < 16) v = expand_bits16<false>(*reinterpret_cast<const uint16_t*>(q_ + (size_t)(qbase + row) * 32 + 2 * kc));
            else if (kc == 16) v.x = 1u;
        }
        const int blk = row >> 7, r = row & 127;
        *reinterpret_cast<uint4*>(sA + blk * TC_TILE_BYTES + (r >> 3) * TC_SBO + kc * TC_LBO + (r & 7) * 16) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (wid == TC_EPI_WARPS) {
        if (lane == 0) {
            const uint8_t* src = texp + (size_t)f0 * tiles_per_frame * TC_TILE_BYTES;
            for (int i = 0; i < ntiles; i++) {
                const int s = i % TC_STAGES;
                mbar_wait(&s_empty[s], ((uint32_t)(i / TC_STAGES) & 1u) ^ 1u);
                mbar_expect_tx(&s_full[s], (uint32_t)TC_TILE_BYTES);
                tma_bulk_g2s(sB + s * TC_TILE_BYTES, src + (size_t)i * TC_TILE_BYTES, (uint32_t)TC_TILE_BYTES, &s_full[s]);
            }
        }
    } else if (wid == TC_EPI_WARPS + 1) {
        if (lane == 0) {
            const uint64_t adesc0 = tc_smem_desc(smem_u32(sA)), adesc1 = tc_smem_desc(smem_u32(sA + TC_TILE_BYTES));
            for (int i = 0; i < ntiles; i++) {
                const int s = i % TC_STAGES, a = i & 1;
                mbar_wait(&s_full[s], (uint32_t)(i / TC_STAGES) & 1u);
                mbar_wait(&s_tempty[a], ((uint32_t)(i >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint64_t bdesc = tc_smem_desc(smem_u32(sB + s * TC_TILE_BYTES));
                const uint32_t d0 = tmem_base + (uint32_t)(a * 256), d1 = d0 + 128u;
#pragma unroll
                for (int k = 0; k < TC_KSTEPS; k++) {
                    tc_mma_i8(d0, adesc0 + (uint64_t)(16 * k), bdesc + (uint64_t)(16 * k), TC_IDESC, k > 0);
                    tc_mma_i8(d1, adesc1 + (uint64_t)(16 * k), bdesc + (uint64_t)(16 * k), TC_IDESC, k > 0);
                }
                tc_commit(&s_empty[s]);
                tc_commit(&s_tfull[a]);
            }
        }
    } else {
        const int blk = (wid >> 2) & 1, half = wid >> 3;
        const int row = (wid & 7) * 32 + lane;                 // this thread's query inside the CTA's 256
        const uint32_t lane_base = (uint32_t)(32 * (wid & 3)) << 16;
        const int T = 128 * thr - 16384;
        int cnt = 0, fl = 0, ti = 0;
        int nt = nf > 0 ? min(counts[f0], cap) : 0;
        for (int i = 0; i < ntiles; i++) {
            const int a = i & 1;
            mbar_wait(&s_tfull[a], (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const int valid = max(min(TC_TN, nt - ti * TC_TN), 0);
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int col0 = half * 64 + c * 32;
                int v[32];
                tc_ld32(tmem_base + lane_base + (uint32_t)(a * 256 + blk * 128 + col0), v);
                if (col0 + 32 <= valid) {
#pragma unroll
                    for (int j = 0; j < 32; j++) cnt += v[j] < T ? 1 : 0;
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) cnt += (col0 + j < valid && v[j] < T) ? 1 : 0;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_tempty[a]);
            if (++ti == tiles_per_frame) {
                // frame f0 + fl is complete for this thread's columns
                int* slot = &s_cnt[fl & 1][row];
                if (cnt) atomicAdd(slot, cnt);
                asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");
                if (half == 0) {
                    int c = qbase + row < nq ? min(*slot, nbest) : 0;
                    *slot = 0;                                  // this buffer is next used two frames on, past the next barrier
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
                    if (lane == 0 && c) atomicAdd(&scores[f0 + fl], c);
                }
                cnt = 0; ti = 0; fl++;
                if (fl < nf) nt = min(counts[f0 + fl], cap);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (wid == TC_EPI_WARPS + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Stored frames [nframes][cap][32] -> [nframes][tiles_per_frame] operand images (rows beyond a frame's count: zeros)
__global__ void __launch_bounds__(256) k_expand_frames(const uint8_t* __restrict__ frames, const int32_t* __restrict__ counts, int cap, int tiles_per_frame,
                                                       uint4* __restrict__ out)
{
    const int f = blockIdx.y;
    const int nt = min(counts[f], cap);
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_frame = (long long)tiles_per_frame * (TC_TILE_BYTES / 16);
    if (u >= per_frame) return;
    const int tile = (int)(u / (TC_TILE_BYTES / 16)), r = (int)(u % (TC_TILE_BYTES / 16));
    const int n1 = r / (TC_KC * 8), kc = (r / 8) % TC_KC, n0 = r & 7;
    const int nl = n1 * 8 + n0, row = tile * TC_TN + nl;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < nt) {
        if (kc < 16) v = expand_bits16<true>(*reinterpret_cast<const uint16_t*>(frames + ((size_t)f * cap + row) * 32 + 2 * kc));
        else if (kc == 16) v.x = (uint32_t)nl;
    }
    out[(size_t)f * per_frame + u] = v;
}

}  // namespace
}  // namespace orbx
