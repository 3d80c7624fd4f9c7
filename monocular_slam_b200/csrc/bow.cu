// bow.cu -- the bag-of-words side of loop closing (the reference vendors DBoW2 for it: ThirdParty/DBoW2), batched on the
// device (sm_100a):
//
//   K14 k_bow_descend   TemplatedVocabulary::transform(feature, word, weight, nid, levelsup) (TemplatedVocabulary.h:1218-1260)
//                       for every descriptor of every frame: at each level the child with the smallest FORB::distance
//                       (FORB.cpp:81-101), the first one on ties.  A group of 16 (or 32) lanes owns a descriptor; each lane
//                       holds one child: the vocabulary is stored so that the children of a node are consecutive 48-byte
//                       records {descriptor, first child, child count, node id, word id}, one dependent load per level.
//   K15 k_bow_build     transform(features, BowVector&, FeatureVector&, levelsup) (:1128-1199): one CTA per frame sorts
//                       (word, feature) and (node, feature) keys in shared memory (std::map order), adds the word weights
//                       as BowVector::addWeight / addIfNotExist do (BowVector.cpp:33-58), normalises as
//                       BowVector::normalize does (:62-87) -- the sum runs in map order on one thread, so the doubles
//                       are the reference's -- and writes the feature vector grouped by node (FeatureVector.cpp:28-43).
//   K16 k_bow_score     GeneralScoring::score of one vector against many (ScoringObject.cpp:24-300): one warp per stored
//                       vector streams its words through a bit table of the query's words, looks the survivors up in the
//                       query, and adds the terms of the common words in ascending word order.
//
// Doubles are IEEE (-fmad=false, correctly rounded division and sqrt), so every output except the KL score (log()) is
// bit-identical to the reference's.
#include <float.h>
#include <math.h>
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace orbx {
namespace {

struct __align__(16) BowNode {      // 48 bytes; the children of a node are consecutive records
    uint32_t desc[8];
    int32_t child_base;             // record index of the first child
    int32_t nchild;
    uint32_t node_id;               // the vocabulary's own node id
    uint32_t word_id;               // word id if the node is a word, else 0 (Node() default in the reference)
};
static_assert(sizeof(BowNode) == 48, "three 16-byte loads per record");

constexpr int BOW_THREADS = 256;

// ---- K14: tree descent.  LPF lanes per feature (16 when no node has more than 16 children).  All 32 lanes of a warp stay in
// the level loop until both of its descriptors have reached a childless node, so every shuffle runs under the full mask
// (xor offsets below LPF stay inside a group).
template <int LPF>
__global__ void __launch_bounds__(BOW_THREADS)
k_bow_descend(const BowNode* __restrict__ nodes, const double* __restrict__ weights, const uint8_t* __restrict__ desc,
              const int32_t* __restrict__ counts, int nframes, int cap, int nid_level,
              uint32_t* __restrict__ word, double* __restrict__ weight, uint32_t* __restrict__ nid)
{
    const int sub = threadIdx.x & (LPF - 1);
    const long long group = ((long long)blockIdx.x * BOW_THREADS + threadIdx.x) / LPF;
    bool active = group < (long long)nframes * cap;
    if (active) {
        const int frame = (int)(group / cap);
        active = (int)(group - (long long)frame * cap) < counts[frame];
    }
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    int base = 0, nch = 0;
    if (active) {
        const uint4* dp = reinterpret_cast<const uint4*>(desc + (size_t)group * 32);
        q0 = __ldg(dp); q1 = __ldg(dp + 1);
        const uint4 root = __ldg(reinterpret_cast<const uint4*>(nodes) + 2);
        base = (int)root.x; nch = (int)root.y;
    }
    uint32_t out_nid = 0, out_word = 0;
    bool have_nid = nid_level <= 0;
    int level = 0, final_rec = 0;
    while (__any_sync(0xffffffffu, nch > 0)) {
        unsigned best = 0xffffffffu;
        for (int c = sub; c < nch; c += LPF) {
            const uint4* rp = reinterpret_cast<const uint4*>(nodes + base + c);
            const uint4 a = __ldg(rp), b = __ldg(rp + 1);
            const unsigned d = __popc(a.x ^ q0.x) + __popc(a.y ^ q0.y) + __popc(a.z ^ q0.z) + __popc(a.w ^ q0.w) +
                               __popc(b.x ^ q1.x) + __popc(b.y ^ q1.y) + __popc(b.z ^ q1.z) + __popc(b.w ^ q1.w);
            best = min(best, (d << 20) | (unsigned)c);           // strict `<` in the reference: the first child wins a tie
        }
#pragma unroll
        for (int o = LPF / 2; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (nch > 0) {
            final_rec = base + (int)(best & 0xfffffu);
            const uint4 info = __ldg(reinterpret_cast<const uint4*>(nodes + final_rec) + 2);    // one address per group; the line is in L1
            ++level;
            if (level == nid_level) { out_nid = info.z; have_nid = true; }
            base = (int)info.x; nch = (int)info.y;
            if (nch == 0) { out_word = info.w; if (!have_nid) out_nid = info.z; }
        }
    }
    if (active && sub == 0) {
        word[group] = out_word;
        weight[group] = level ? weights[final_rec] : 0.0;
        nid[group] = out_nid;
    }
}

// ---- K15: one frame's BowVector (blockIdx.y == 0) or FeatureVector (blockIdx.y == 1)
constexpr int BOW_BUILD_THREADS = 1024;

template <typename KeyT>
__device__ __forceinline__ void bitonic_sort(KeyT* keys, int P)
{
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < P / 2; t += BOW_BUILD_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), ixj = i | j;      // the t-th pair of this stage
                const KeyT a = keys[i], b = keys[ixj];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
            }
            __syncthreads();
        }
}

// pos[i] = number of run heads before i (i is a head when its key's id differs from its predecessor's); returns the number of runs
template <typename KeyT>
__device__ int run_offsets(const KeyT* keys, int n, int shift, int* pos, int* s_warp, int* s_carry)
{
    if (threadIdx.x == 0) *s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i0 = 0; i0 < n; i0 += BOW_BUILD_THREADS) {
        const int i = i0 + threadIdx.x;
        const bool head = i < n && (i == 0 || (keys[i] >> shift) != (keys[i - 1] >> shift));
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {                       // exclusive scan of the 32 warp totals
            const int v = s_warp[lane];
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
            s_warp[lane] = incl - v + *s_carry;
            __syncwarp();
            if (lane == 31) *s_carry += incl;
        }
        __syncthreads();
        if (i < n) pos[i] = s_warp[warp] + __popc(bal & ((1u << lane) - 1));
        __syncthreads();
    }
    return *s_carry;
}

template <typename KeyT>
__global__ void __launch_bounds__(BOW_BUILD_THREADS)
k_bow_build(const uint32_t* __restrict__ word, const double* __restrict__ weight, const uint32_t* __restrict__ nid,
            const int32_t* __restrict__ counts, int cap, int P, int shift, int accumulate, int must, int l2,
            uint32_t* __restrict__ bow_words, double* bow_vals, int32_t* __restrict__ nbow,
            uint32_t* __restrict__ fv_nodes, int32_t* __restrict__ fv_offsets, uint32_t* __restrict__ fv_feats, int32_t* __restrict__ nfv,
            int first_part)
{
    extern __shared__ __align__(16) unsigned char bow_smem[];
    KeyT* keys = reinterpret_cast<KeyT*>(bow_smem);                                 // [P]
    int* pos = reinterpret_cast<int*>(bow_smem + (size_t)P * sizeof(KeyT));         // [P]
    __shared__ int s_warp[32];
    __shared__ int s_carry, s_nvalid;
    __shared__ double s_norm;
    const int frame = blockIdx.x, n = min(counts[frame], cap);
    const bool feature_vector = first_part + (int)blockIdx.y == 1;
    const size_t f0 = (size_t)frame * cap;
    const uint32_t* id = (feature_vector ? nid : word) + f0;
    weight += f0;
    const KeyT idx_mask = (KeyT)(((KeyT)1 << shift) - 1);

    // (id, feature) keys of the features whose word is not stopped (w > 0); the others sort to the end
    if (threadIdx.x == 0) s_nvalid = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < P; i += BOW_BUILD_THREADS) {
        const bool ok = i < n && weight[i] > 0;
        keys[i] = ok ? (KeyT)(((KeyT)id[i] << shift) | (KeyT)i) : (KeyT)~(KeyT)0;
        mine += ok;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_nvalid, mine);
    __syncthreads();
    const int nvalid = s_nvalid;
    bitonic_sort(keys, P);
    const int nruns = run_offsets(keys, nvalid, shift, pos, s_warp, &s_carry);

    if (feature_vector) {
        // nodes ascending, each node's features in the order they were added (FeatureVector::addFeature)
        fv_nodes += f0; fv_feats += f0; fv_offsets += (size_t)frame * (cap + 1);
        for (int i = threadIdx.x; i < nvalid; i += BOW_BUILD_THREADS) {
            fv_feats[i] = (uint32_t)(keys[i] & idx_mask);
            if (i == 0 || (keys[i - 1] >> shift) != (keys[i] >> shift)) { fv_nodes[pos[i]] = (uint32_t)(keys[i] >> shift); fv_offsets[pos[i]] = i; }
        }
        if (threadIdx.x == 0) { fv_offsets[nruns] = nvalid; nfv[frame] = nruns; }
        return;
    }

    bow_words += f0; bow_vals += f0;
    const int nb = nruns;
    for (int i = threadIdx.x; i < nvalid; i += BOW_BUILD_THREADS) {
        const uint32_t w = (uint32_t)(keys[i] >> shift);
        if (i == 0 || (uint32_t)(keys[i - 1] >> shift) != w) {
            const double wt = weight[(uint32_t)(keys[i] & idx_mask)];
            double v = wt;
            if (accumulate)                                   // addWeight: v += w once per further occurrence, in order
                for (int j = i + 1; j < nvalid && (uint32_t)(keys[j] >> shift) == w; j++) v += wt;
            bow_words[pos[i]] = w;
            bow_vals[pos[i]] = v;
        }
    }
    if (threadIdx.x == 0) nbow[frame] = nb;
    if (!(accumulate && nb > 0 && !must) && !must) return;
    __syncthreads();
    if (!must) {                                              // TF / TF_IDF without normalisation: v /= v.size()
        const double nd = (double)nb;
        for (int i = threadIdx.x; i < nb; i += BOW_BUILD_THREADS) bow_vals[i] /= nd;
        return;
    }
    // BowVector::normalize sums in map order: the values go to shared memory (the keys are dead), one thread adds them up
    double* sv = reinterpret_cast<double*>(bow_smem);
    for (int i = threadIdx.x; i < nb; i += BOW_BUILD_THREADS) sv[i] = bow_vals[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        double norm = 0.0;
        if (!l2) for (int i = 0; i < nb; i++) norm += fabs(sv[i]);
        else { for (int i = 0; i < nb; i++) norm += sv[i] * sv[i]; norm = sqrt(norm); }
        s_norm = norm;
    }
    __syncthreads();
    const double norm = s_norm;
    if (norm > 0.0)
        for (int i = threadIdx.x; i < nb; i += BOW_BUILD_THREADS) bow_vals[i] = sv[i] / norm;
}

// ---- K16: one query vector against M stored vectors
enum { S_L1 = 0, S_L2 = 1, S_CHI = 2, S_KL = 3, S_BHAT = 4, S_DOT = 5 };

__device__ __forceinline__ int lower_bound_u32(const uint32_t* a, int n, uint32_t key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// The query's words sit in shared memory twice: sorted (for the position of a common word) and as a 2^17-bit table indexed by
// the low bits of the word id, which answers "not in the query" for ~98 % of a stored vector's words with one load.  A warp
// walks its stored vectors 128 words at a time (four coalesced loads in flight) and adds the terms of the common words in
// ascending word order, as the reference's merge does.
constexpr int BOW_FILTER_BITS = 17;
constexpr int BOW_FILTER_WORDS = 1 << (BOW_FILTER_BITS - 5);
constexpr int BOW_BUCKETS = 1024;
constexpr int BOW_UNROLL = 4;            // 32-word loads a warp keeps in flight
constexpr int BOW_QUEUE = 32 + 32 * BOW_UNROLL;      // a queue is drained at 32 entries; 32 * BOW_UNROLL words can arrive before the next check

// the query's bit table and bucket starts, built once per call in global memory (every CTA of k_bow_score copies them)
__global__ void __launch_bounds__(1024)
k_bow_query_tables(const uint32_t* __restrict__ qwords, int nq, int word_shift, uint32_t* __restrict__ tables, int* __restrict__ next_entry)
{
    if (threadIdx.x == 0) *next_entry = 0;        // k_bow_score's work counter
    uint32_t* filter = tables;                                                   // [BOW_FILTER_WORDS]
    uint16_t* bucket = reinterpret_cast<uint16_t*>(tables + BOW_FILTER_WORDS);   // [BOW_BUCKETS + 1] (+ padding)
    for (int i = threadIdx.x; i < BOW_FILTER_WORDS; i += blockDim.x) filter[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i <= nq; i += blockDim.x) {
        const int prev = i ? min((int)(qwords[i - 1] >> word_shift), BOW_BUCKETS - 1) : -1;
        int cur = BOW_BUCKETS;
        if (i < nq) {
            const uint32_t w = qwords[i];
            atomicOr(&filter[(w >> 5) & (BOW_FILTER_WORDS - 1)], 1u << (w & 31));
            cur = min((int)(w >> word_shift), BOW_BUCKETS - 1);
        }
        for (int bk = prev + 1; bk <= cur; bk++) bucket[bk] = (uint16_t)i;       // ascending words: every bucket is written once
    }
}
constexpr int BOW_TABLE_WORDS = BOW_FILTER_WORDS + (BOW_BUCKETS + 8) / 2;       // uint32 words of the two tables, a multiple of 4

template <int SCORING>
__global__ void __launch_bounds__(BOW_THREADS)
k_bow_score(const uint32_t* __restrict__ qwords, const double* __restrict__ qvals, int nq, const uint32_t* __restrict__ tables,
            const int64_t* __restrict__ db_start, const int32_t* __restrict__ db_count, const uint32_t* __restrict__ db_words,
            const double* __restrict__ db_vals, int M, int word_shift, int vals_in_smem, int* __restrict__ next_entry,
            double* __restrict__ scores)
{
    extern __shared__ __align__(16) unsigned char bow_smem[];
    uint32_t* s_filter = reinterpret_cast<uint32_t*>(bow_smem);                     // [BOW_FILTER_WORDS]
    const uint16_t* s_bucket = reinterpret_cast<const uint16_t*>(s_filter + BOW_FILTER_WORDS);   // first query word of every range of 2^word_shift ids
    double* s_qv = vals_in_smem ? reinterpret_cast<double*>(s_filter + BOW_TABLE_WORDS) : nullptr;       // the query's values, if they fit
    uint32_t* s_q = s_filter + BOW_TABLE_WORDS + (vals_in_smem ? 2 * nq : 0);                            // the query's words, ascending
    __shared__ uint32_t s_queue[BOW_THREADS / 32][BOW_QUEUE];    // per warp: positions in the stored vector of the words that passed the table
    __shared__ double s_terms[BOW_THREADS / 32][32];             // per warp: the terms of 32 queue entries
    static_assert(BOW_TABLE_WORDS % 4 == 0 && BOW_FILTER_WORDS % 4 == 0, "the tables are copied as 16-byte words");
    for (int i = threadIdx.x; i < BOW_TABLE_WORDS / 4; i += blockDim.x)
        reinterpret_cast<uint4*>(s_filter)[i] = __ldg(reinterpret_cast<const uint4*>(tables) + i);
    for (int i = threadIdx.x; i < nq; i += blockDim.x) {
        s_q[i] = qwords[i];
        if (s_qv) s_qv[i] = qvals[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // stored vectors are handed out one at a time: a vector that shares many words with the query takes much longer than one
    // that shares few, and a fixed assignment would let two of those land on one warp
    for (;;) {
        int e = 0;
        if (lane == 0) e = atomicAdd(next_entry, 1);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= M) break;
        const uint32_t* w2 = db_words + db_start[e];
        const double* v2 = db_vals + db_start[e];
        const int n2 = db_count[e];
        double score = 0;
        if (SCORING == S_KL) {
            // every word of the query contributes, in ascending order: with its partner in the stored vector, or alone
            const double log_eps = log(DBL_EPSILON);
            const uint32_t last2 = n2 > 0 ? w2[n2 - 1] : 0;
            for (int i0 = 0; i0 < nq; i0 += 32) {
                const int i = i0 + lane;
                double term = 0;
                bool add = false;
                if (i < nq) {
                    const uint32_t w = s_q[i];
                    const double vi = qvals[i];
                    const int j = lower_bound_u32(w2, n2, w);
                    if (j < n2 && w2[j] == w) { const double wi = v2[j]; if (vi != 0 && wi != 0) { term = vi * log(vi / wi); add = true; } }
                    else if (n2 > 0 && w < last2) { term = vi * (log(vi) - log_eps); add = true; }       // inside the merge loop: unconditional
                    else if (vi != 0) { term = vi * (log(vi) - log_eps); add = true; }                   // after it: zero values skipped
                }
                unsigned bal = __ballot_sync(0xffffffffu, add);
                while (bal) {
                    const int b = __ffs(bal) - 1;
                    bal &= bal - 1;
                    score += __shfl_sync(0xffffffffu, term, b);
                }
            }
        } else {
            // words that pass the table wait, in order, in a per-warp queue; a full queue is resolved 32 at a time: exact
            // position in the query, the term, and the terms of the common words added in lane (= word) order
            uint32_t* queue = s_queue[threadIdx.x >> 5];
            double* terms = s_terms[threadIdx.x >> 5];
            // look(t): is queue entry t a word of the query, and if so the two values (loads issued, not yet used)
            auto look = [&](int t, int count, double& vi, double& wi) -> bool {
                if (t >= count) return false;
                const uint32_t j = queue[t], w = __ldg(w2 + j);       // the word again: it is in L1 / L2
                const int bk = min((int)(w >> word_shift), BOW_BUCKETS - 1);
                const int lo = s_bucket[bk];
                const int i = lo + lower_bound_u32(s_q + lo, s_bucket[bk + 1] - lo, w);
                if (i >= nq || s_q[i] != w) return false;
                vi = s_qv ? s_qv[i] : qvals[i];
                wi = v2[j];
                return true;
            };
            auto drain = [&](int count) {
                double vi = 0, wi = 0;
                bool hit = look(lane, count, vi, wi);
                for (int g = 0; g < count; g += 32) {
                    double vn = 0, wn_ = 0;
                    const bool hit_next = look(g + 32 + lane, count, vn, wn_);      // the next 32 are fetched while these are added
                    double term = 0;
                    if (hit) {
                        if (SCORING == S_L1) term = fabs(vi - wi) - fabs(vi) - fabs(wi);
                        else if (SCORING == S_L2 || SCORING == S_DOT) term = vi * wi;
                        else if (SCORING == S_CHI) { hit = vi + wi != 0.0; if (hit) term = vi * wi / (vi + wi); }
                        else term = sqrt(vi * wi);
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, hit);
                    if (bal) {
                        // through shared memory, not shuffles: the 32 loads are independent of the sum, so the chain of
                        // additions runs at the latency of an addition
                        terms[lane] = term;
                        __syncwarp();
#pragma unroll
                        for (int b = 0; b < 32; b++)
                            if ((bal >> b) & 1) score += terms[b];
                        __syncwarp();
                    }
                    hit = hit_next; vi = vn; wi = wn_;
                }
            };
            int qn = 0;
            uint32_t w[BOW_UNROLL], wn[BOW_UNROLL];
#pragma unroll
            for (int u = 0; u < BOW_UNROLL; u++) { const int j = u * 32 + lane; wn[u] = j < n2 ? __ldg(w2 + j) : 0xffffffffu; }
            for (int j0 = 0; j0 < n2; j0 += 32 * BOW_UNROLL) {
#pragma unroll
                for (int u = 0; u < BOW_UNROLL; u++) {   // the next 32 * BOW_UNROLL words are on their way while these are looked at
                    w[u] = wn[u];
                    const int j = j0 + 32 * BOW_UNROLL + u * 32 + lane;
                    wn[u] = j < n2 ? __ldg(w2 + j) : 0xffffffffu;
                }
#pragma unroll
                for (int u = 0; u < BOW_UNROLL; u++) {
                    const int j = j0 + u * 32 + lane;
                    const bool maybe = j < n2 && ((s_filter[(w[u] >> 5) & (BOW_FILTER_WORDS - 1)] >> (w[u] & 31)) & 1);
                    const unsigned bal = __ballot_sync(0xffffffffu, maybe);
                    if (maybe) {
                        queue[qn + __popc(bal & ((1u << lane) - 1))] = (uint32_t)j;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(v2 + j));      // its value is wanted when the queue is drained
                    }
                    qn += __popc(bal);
                }
                if (qn >= 32) { __syncwarp(); drain(qn); qn = 0; __syncwarp(); }
            }
            if (qn) { __syncwarp(); drain(qn); }
            __syncwarp();
        }
        if (SCORING == S_L1) score = -score / 2.0;
        else if (SCORING == S_L2) score = score >= 1 ? 1.0 : 1.0 - sqrt(1.0 - score);
        else if (SCORING == S_CHI) score = 2. * score;
        if (lane == 0) scores[e] = score;
    }
}

template <typename T>
int bow_grow(T** p, size_t* have, size_t want)
{
    if (*p && *have >= want) return ORBX_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *have = 0; }
    const size_t bytes = align_up(want + want / 4, 256);
    cudaError_t e = cudaMalloc((void**)p, bytes);
    if (e != cudaSuccess) { *p = nullptr; set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return ORBX_E_ALLOC; }
    *have = bytes;
    return ORBX_OK;
}

}  // namespace
}  // namespace orbx

using namespace orbx;

struct bowx_context {
    int device;
    cudaStream_t own_stream, stream;
    size_t smem_optin;
    int sm_count;
    // vocabulary: host copy in the reference's numbering, device copy in child-contiguous record order
    int k, L, scoring, weighting, nnodes, nwords, max_children;
    std::vector<int32_t> parent, word_node, rec_of_node;
    std::vector<double> weight;          // by node id
    BowNode* d_nodes; size_t nodes_bytes;
    double* d_weights; size_t weights_bytes;     // by record
    // per-feature results of the descent (the staging of the host forms follows in d_buf)
    uint32_t* d_word; size_t word_bytes;
    double* d_weight; size_t weight_bytes;
    uint32_t* d_nid; size_t nid_bytes;
    uint8_t* d_buf; size_t buf_bytes;
    int* d_next;                         // k_bow_score's work counter
    uint32_t* d_tables;                  // the query's bit table and bucket starts (k_bow_query_tables)
};

extern "C" int bowx_destroy(bowx_handle h)
{
    if (!h) return ORBX_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_nodes); cudaFree(h->d_weights); cudaFree(h->d_word); cudaFree(h->d_weight); cudaFree(h->d_nid); cudaFree(h->d_buf); cudaFree(h->d_next); cudaFree(h->d_tables);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return ORBX_OK;
}

extern "C" int bowx_create(bowx_handle* out, int device)
{
    ORBX_REQUIRE(out != nullptr, "bowx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { set_error("bowx_create: no CUDA device (%s); liborbx has no CPU fallback", cudaGetErrorString(e)); return ORBX_E_CUDA; }
    ORBX_REQUIRE(device >= 0 && device < ndev, "bowx_create: device %d out of range [0,%d)", device, ndev);
    ORBX_CUDA(cudaSetDevice(device));
    bowx_context* h = new bowx_context();
    h->device = device;
    h->own_stream = h->stream = nullptr;
    h->k = h->L = h->scoring = h->weighting = h->nnodes = h->nwords = h->max_children = 0;
    h->d_next = nullptr; h->d_tables = nullptr;
    h->d_nodes = nullptr; h->d_weights = nullptr; h->d_word = nullptr; h->d_weight = nullptr; h->d_nid = nullptr; h->d_buf = nullptr;
    h->nodes_bytes = h->weights_bytes = h->word_bytes = h->weight_bytes = h->nid_bytes = h->buf_bytes = 0;
    ORBX_CUDA_OR(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking), bowx_destroy(h));
    h->stream = h->own_stream;
    int optin = 0;
    ORBX_CUDA_OR(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device), bowx_destroy(h));
    ORBX_CUDA_OR(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device), bowx_destroy(h));
    ORBX_CUDA_OR(cudaMalloc((void**)&h->d_next, sizeof(int)), bowx_destroy(h));
    ORBX_CUDA_OR(cudaMalloc((void**)&h->d_tables, sizeof(uint32_t) * BOW_TABLE_WORDS), bowx_destroy(h));
    h->smem_optin = (size_t)optin - 1024;       // dynamic part: k_bow_build also has a few static words
    ORBX_CUDA_OR(cudaFuncSetAttribute(k_bow_score<S_L1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), bowx_destroy(h));
    ORBX_CUDA_OR(cudaFuncSetAttribute(k_bow_score<S_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), bowx_destroy(h));
    ORBX_CUDA_OR(cudaFuncSetAttribute(k_bow_score<S_CHI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), bowx_destroy(h));
    ORBX_CUDA_OR(cudaFuncSetAttribute(k_bow_score<S_KL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), bowx_destroy(h));
    ORBX_CUDA_OR(cudaFuncSetAttribute(k_bow_score<S_BHAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), bowx_destroy(h));
    ORBX_CUDA_OR(cudaFuncSetAttribute(k_bow_score<S_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), bowx_destroy(h));
    ORBX_CUDA_OR(cudaFuncSetAttribute(k_bow_build<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024), bowx_destroy(h));
    ORBX_CUDA_OR(cudaFuncSetAttribute(k_bow_build<unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024), bowx_destroy(h));
    *out = h;
    return ORBX_OK;
}

extern "C" int bowx_set_stream(bowx_handle h, void* cuda_stream)
{
    ORBX_REQUIRE(h != nullptr, "bowx_set_stream: NULL handle");
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return ORBX_OK;
}

extern "C" int bowx_get_stream(bowx_handle h, void** cuda_stream)
{
    ORBX_REQUIRE(h != nullptr && cuda_stream != nullptr, "bowx_get_stream: NULL argument");
    *cuda_stream = h->stream == h->own_stream ? nullptr : (void*)h->stream;
    return ORBX_OK;
}

extern "C" int bowx_synchronize(bowx_handle h)
{
    ORBX_REQUIRE(h != nullptr, "bowx_synchronize: NULL handle");
    ORBX_CUDA(cudaSetDevice(h->device));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

static int upload_weights(bowx_handle h)
{
    std::vector<double> by_rec((size_t)h->nnodes);
    for (int i = 0; i < h->nnodes; i++) by_rec[(size_t)h->rec_of_node[i]] = h->weight[(size_t)i];
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    ORBX_CUDA(cudaMemcpy(h->d_weights, by_rec.data(), sizeof(double) * (size_t)h->nnodes, cudaMemcpyHostToDevice));
    return ORBX_OK;
}

static int set_vocabulary(bowx_handle h, int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent, const uint8_t* leaf,
                          const uint8_t* desc, const double* weight);

extern "C" int bowx_set_vocabulary(bowx_handle h, int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent,
                                   const uint8_t* leaf, const uint8_t* desc, const double* weight)
{
    ORBX_NOTHROW(set_vocabulary(h, k, L, scoring, weighting, nnodes, parent, leaf, desc, weight))
}

static int set_vocabulary(bowx_handle h, int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent, const uint8_t* leaf,
                          const uint8_t* desc, const double* weight)
{
    ORBX_REQUIRE(h != nullptr, "bowx_set_vocabulary: NULL handle");
    ORBX_REQUIRE(parent && leaf && desc && weight, "bowx_set_vocabulary: NULL pointer");
    ORBX_REQUIRE(nnodes >= 1 && nnodes < (1 << 20) * 16, "bowx_set_vocabulary: %d nodes out of range", nnodes);
    ORBX_REQUIRE(scoring >= BOWX_L1_NORM && scoring <= BOWX_DOT_PRODUCT, "bowx_set_vocabulary: scoring type %d", scoring);
    ORBX_REQUIRE(weighting >= BOWX_TF_IDF && weighting <= BOWX_BINARY, "bowx_set_vocabulary: weighting type %d", weighting);
    for (int i = 1; i < nnodes; i++)
        ORBX_REQUIRE(parent[i] >= 0 && parent[i] < i, "bowx_set_vocabulary: parent[%d] = %d is not an earlier node", i, parent[i]);
    ORBX_CUDA(cudaSetDevice(h->device));
    // children of every node in node order (loadFromTextFile appends them as it reads)
    std::vector<int32_t> off((size_t)nnodes + 1, 0), child((size_t)nnodes);
    for (int i = 1; i < nnodes; i++) off[(size_t)parent[i] + 1]++;
    int max_children = 0;
    for (int i = 0; i < nnodes; i++) { max_children = std::max(max_children, off[(size_t)i + 1]); off[(size_t)i + 1] += off[(size_t)i]; }
    ORBX_REQUIRE(max_children < (1 << 20), "bowx_set_vocabulary: a node with %d children", max_children);
    {
        std::vector<int32_t> fill(off.begin(), off.end() - 1);
        for (int i = 1; i < nnodes; i++) child[(size_t)fill[(size_t)parent[i]]++] = i;
    }
    // records: breadth first, every node's children consecutive
    std::vector<BowNode> rec((size_t)nnodes);
    std::vector<int32_t> rec_of_node((size_t)nnodes, -1), order;
    order.reserve((size_t)nnodes);
    order.push_back(0);
    rec_of_node[0] = 0;
    int next = 1;
    for (size_t q = 0; q < order.size(); q++) {
        const int node = order[q];
        for (int c = off[(size_t)node]; c < off[(size_t)node + 1]; c++) { rec_of_node[(size_t)child[(size_t)c]] = next++; order.push_back(child[(size_t)c]); }
    }
    h->word_node.clear();
    std::vector<uint32_t> word_id((size_t)nnodes, 0);
    for (int i = 1; i < nnodes; i++)
        if (leaf[i]) { word_id[(size_t)i] = (uint32_t)h->word_node.size(); h->word_node.push_back(i); }
    for (int i = 0; i < nnodes; i++) {
        BowNode& r = rec[(size_t)rec_of_node[(size_t)i]];
        memcpy(r.desc, desc + (size_t)i * 32, 32);
        const int nch = off[(size_t)i + 1] - off[(size_t)i];
        r.nchild = nch;
        r.child_base = nch ? rec_of_node[(size_t)child[(size_t)off[(size_t)i]]] : 0;
        r.node_id = (uint32_t)i;
        r.word_id = word_id[(size_t)i];
    }
    int rc = bow_grow(&h->d_nodes, &h->nodes_bytes, sizeof(BowNode) * (size_t)nnodes);
    if (!rc) rc = bow_grow(&h->d_weights, &h->weights_bytes, sizeof(double) * (size_t)nnodes);
    if (rc) return rc;
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    ORBX_CUDA(cudaMemcpy(h->d_nodes, rec.data(), sizeof(BowNode) * (size_t)nnodes, cudaMemcpyHostToDevice));
    h->k = k; h->L = L; h->scoring = scoring; h->weighting = weighting; h->nnodes = nnodes; h->nwords = (int)h->word_node.size();
    h->max_children = max_children;
    h->parent.assign(parent, parent + nnodes);
    h->parent[0] = 0;
    h->weight.assign(weight, weight + nnodes);
    h->rec_of_node.swap(rec_of_node);
    return upload_weights(h);
}

extern "C" int bowx_vocabulary_info(bowx_handle h, int32_t* info)
{
    ORBX_REQUIRE(h != nullptr && info != nullptr, "bowx_vocabulary_info: NULL argument");
    info[0] = h->k; info[1] = h->L; info[2] = h->scoring; info[3] = h->weighting; info[4] = h->nnodes; info[5] = h->nwords;
    return ORBX_OK;
}

extern "C" int bowx_stop_words(bowx_handle h, double min_weight, int32_t* count)
{
    ORBX_REQUIRE(h != nullptr && count != nullptr, "bowx_stop_words: NULL argument");
    ORBX_REQUIRE(h->nnodes > 0, "bowx_stop_words: no vocabulary");
    ORBX_CUDA(cudaSetDevice(h->device));
    int c = 0;
    for (size_t w = 0; w < h->word_node.size(); w++) {
        double& wt = h->weight[(size_t)h->word_node[w]];
        if (wt < min_weight) { c++; wt = 0; }
    }
    *count = c;
    return upload_weights(h);
}

extern "C" int bowx_parent_node(bowx_handle h, uint32_t word, int levelsup, uint32_t* node)
{
    ORBX_REQUIRE(h != nullptr && node != nullptr, "bowx_parent_node: NULL argument");
    ORBX_REQUIRE(word < (uint32_t)h->nwords, "bowx_parent_node: word %u of %d", word, h->nwords);
    int ret = h->word_node[word];
    while (levelsup > 0 && ret != 0) { --levelsup; ret = h->parent[(size_t)ret]; }
    *node = (uint32_t)ret;
    return ORBX_OK;
}

extern "C" int bowx_word_weight(bowx_handle h, uint32_t word, double* weight)
{
    ORBX_REQUIRE(h != nullptr && weight != nullptr, "bowx_word_weight: NULL argument");
    ORBX_REQUIRE(word < (uint32_t)h->nwords, "bowx_word_weight: word %u of %d", word, h->nwords);
    *weight = h->weight[(size_t)h->word_node[word]];
    return ORBX_OK;
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

static int descend(bowx_handle h, const uint8_t* d_desc, const int32_t* d_counts, int nframes, int cap, int levelsup,
                   uint32_t* d_word, double* d_weight, uint32_t* d_nid)
{
    const long long total = (long long)nframes * cap;
    if (total == 0) return ORBX_OK;
    const int nid_level = h->L - levelsup;
    if (h->max_children <= 16) {
        const long long blocks = (total * 16 + BOW_THREADS - 1) / BOW_THREADS;
        k_bow_descend<16><<<(unsigned)blocks, BOW_THREADS, 0, h->stream>>>(h->d_nodes, h->d_weights, d_desc, d_counts, nframes, cap, nid_level,
                                                                         d_word, d_weight, d_nid);
    } else {
        const long long blocks = (total * 32 + BOW_THREADS - 1) / BOW_THREADS;
        k_bow_descend<32><<<(unsigned)blocks, BOW_THREADS, 0, h->stream>>>(h->d_nodes, h->d_weights, d_desc, d_counts, nframes, cap, nid_level,
                                                                         d_word, d_weight, d_nid);
    }
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

static int check_transform_args(bowx_handle h, const char* fn, int nframes, int cap)
{
    ORBX_REQUIRE(h != nullptr, "%s: NULL handle", fn);
    ORBX_REQUIRE(h->nnodes > 0, "%s: no vocabulary (bowx_set_vocabulary)", fn);
    ORBX_REQUIRE(nframes >= 0 && cap >= 1, "%s: nframes %d / cap %d out of range", fn, nframes, cap);
    ORBX_REQUIRE((long long)nframes * cap * 32 < (1LL << 40), "%s: %d frames x %d features is too large", fn, nframes, cap);
    return ORBX_OK;
}

extern "C" int bowx_transform_features_dev(bowx_handle h, const uint8_t* d_desc, const int32_t* d_counts, int nframes, int cap, int levelsup,
                                           uint32_t* d_word, double* d_weight, uint32_t* d_nid)
{
    int rc = check_transform_args(h, "bowx_transform_features_dev", nframes, cap);
    if (rc) return rc;
    ORBX_REQUIRE(d_desc && d_counts && d_word && d_weight && d_nid, "bowx_transform_features_dev: NULL pointer");
    ORBX_CUDA(cudaSetDevice(h->device));
    return descend(h, d_desc, d_counts, nframes, cap, levelsup, d_word, d_weight, d_nid);
}

extern "C" int bowx_transform_batch_dev(bowx_handle h, const uint8_t* d_desc, const int32_t* d_counts, int nframes, int cap, int levelsup,
                                        uint32_t* d_bow_words, double* d_bow_vals, int32_t* d_nbow, uint32_t* d_fv_nodes,
                                        int32_t* d_fv_offsets, uint32_t* d_fv_feats, int32_t* d_nfv)
{
    int rc = check_transform_args(h, "bowx_transform_batch_dev", nframes, cap);
    if (rc) return rc;
    ORBX_REQUIRE(d_desc && d_counts && d_bow_words && d_bow_vals && d_nbow, "bowx_transform_batch_dev: NULL pointer");
    const bool fv = d_fv_nodes || d_fv_offsets || d_fv_feats || d_nfv;
    ORBX_REQUIRE(!fv || (d_fv_nodes && d_fv_offsets && d_fv_feats && d_nfv), "bowx_transform_batch_dev: the feature vector needs all four of its arrays");
    // keys are (id << idx_bits) | feature index: 32 bits when the ids leave room for the index, else 64
    const int P = next_pow2(std::max(cap, 64));
    int idx_bits = 0;
    while ((1 << idx_bits) < cap) idx_bits++;
    const bool bow32 = idx_bits < 32 && (unsigned long long)std::max(h->nwords, 1) < (1ull << (32 - idx_bits)) - 1;
    const bool fv32 = idx_bits < 32 && (unsigned long long)h->nnodes < (1ull << (32 - idx_bits)) - 1;
    const size_t smem64 = (size_t)P * (sizeof(unsigned long long) + sizeof(int));
    ORBX_REQUIRE(smem64 <= h->smem_optin, "bowx_transform_batch_dev: cap %d needs %zu bytes of shared memory (limit %zu): at most 16384 features per frame", cap, smem64, h->smem_optin);
    if (nframes == 0) return ORBX_OK;
    ORBX_CUDA(cudaSetDevice(h->device));
    const size_t total = (size_t)nframes * cap;
    rc = bow_grow(&h->d_word, &h->word_bytes, total * sizeof(uint32_t));
    if (!rc) rc = bow_grow(&h->d_weight, &h->weight_bytes, total * sizeof(double));
    if (!rc) rc = bow_grow(&h->d_nid, &h->nid_bytes, total * sizeof(uint32_t));
    if (rc) return rc;
    if (h->nwords == 0) {          // empty(): transform() returns empty vectors
        ORBX_CUDA(cudaMemsetAsync(d_nbow, 0, sizeof(int32_t) * (size_t)nframes, h->stream));
        if (fv) {
            ORBX_CUDA(cudaMemsetAsync(d_nfv, 0, sizeof(int32_t) * (size_t)nframes, h->stream));
            ORBX_CUDA(cudaMemsetAsync(d_fv_offsets, 0, sizeof(int32_t) * (size_t)nframes * ((size_t)cap + 1), h->stream));
        }
        return ORBX_OK;
    }
    rc = descend(h, d_desc, d_counts, nframes, cap, levelsup, h->d_word, h->d_weight, h->d_nid);
    if (rc) return rc;
    const int accumulate = h->weighting == BOWX_TF || h->weighting == BOWX_TF_IDF;
    const int must = h->scoring != BOWX_DOT_PRODUCT, l2 = h->scoring == BOWX_L2_NORM;
    // one CTA per frame and vector; both vectors in one launch when their keys have the same width
    const size_t smem32 = (size_t)P * 8;        // 4-byte keys + offsets; also holds the P doubles of the normalisation
#define BOW_BUILD(KEY, SMEM, SHIFT, PARTS, FIRST)                                                                                   \
    k_bow_build<KEY><<<dim3((unsigned)nframes, PARTS), BOW_BUILD_THREADS, SMEM, h->stream>>>(                                      \
        h->d_word, h->d_weight, h->d_nid, d_counts, cap, P, SHIFT, accumulate, must, l2, d_bow_words, d_bow_vals, d_nbow, d_fv_nodes, \
        d_fv_offsets, d_fv_feats, d_nfv, FIRST)
    if (!fv) {
        if (bow32) BOW_BUILD(uint32_t, smem32, idx_bits, 1, 0); else BOW_BUILD(unsigned long long, smem64, 32, 1, 0);
    } else if (bow32 == fv32) {
        if (bow32) BOW_BUILD(uint32_t, smem32, idx_bits, 2, 0); else BOW_BUILD(unsigned long long, smem64, 32, 2, 0);
    } else {
        if (bow32) BOW_BUILD(uint32_t, smem32, idx_bits, 1, 0); else BOW_BUILD(unsigned long long, smem64, 32, 1, 0);
        ORBX_CUDA(cudaGetLastError());
        if (fv32) BOW_BUILD(uint32_t, smem32, idx_bits, 1, 1); else BOW_BUILD(unsigned long long, smem64, 32, 1, 1);
    }
#undef BOW_BUILD
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

// carve `bytes` (rounded to 256) out of the staging buffer
static uint8_t* carve(uint8_t*& p, size_t bytes) { uint8_t* r = p; p += align_up(bytes, 256); return r; }

extern "C" int bowx_transform_features(bowx_handle h, const uint8_t* desc, int n, int levelsup, uint32_t* word, double* weight, uint32_t* node)
{
    int rc = check_transform_args(h, "bowx_transform_features", 1, std::max(n, 1));
    if (rc) return rc;
    ORBX_REQUIRE(n >= 0, "bowx_transform_features: n %d", n);
    if (n == 0) return ORBX_OK;
    ORBX_REQUIRE(desc && word && weight && node, "bowx_transform_features: NULL pointer");
    ORBX_CUDA(cudaSetDevice(h->device));
    const size_t need = align_up((size_t)n * 32, 256) + 256 + align_up((size_t)n * 4, 256) * 2 + align_up((size_t)n * 8, 256);
    rc = bow_grow(&h->d_buf, &h->buf_bytes, need);
    if (rc) return rc;
    uint8_t* p = h->d_buf;
    uint8_t* d_desc = carve(p, (size_t)n * 32);
    int32_t* d_count = (int32_t*)carve(p, 4);
    uint32_t* d_word = (uint32_t*)carve(p, (size_t)n * 4);
    uint32_t* d_nid = (uint32_t*)carve(p, (size_t)n * 4);
    double* d_weight = (double*)carve(p, (size_t)n * 8);
    ORBX_CUDA(cudaMemcpyAsync(d_desc, desc, (size_t)n * 32, cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(d_count, &n, 4, cudaMemcpyHostToDevice, h->stream));
    rc = descend(h, d_desc, d_count, 1, n, levelsup, d_word, d_weight, d_nid);
    if (rc) return rc;
    ORBX_CUDA(cudaMemcpyAsync(word, d_word, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(weight, d_weight, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(node, d_nid, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int bowx_transform_batch(bowx_handle h, const uint8_t* desc, const int32_t* counts, int nframes, int cap, int levelsup,
                                    uint32_t* bow_words, double* bow_vals, int32_t* nbow, uint32_t* fv_nodes, int32_t* fv_offsets,
                                    uint32_t* fv_feats, int32_t* nfv)
{
    int rc = check_transform_args(h, "bowx_transform_batch", nframes, cap);
    if (rc) return rc;
    if (nframes == 0) return ORBX_OK;
    ORBX_REQUIRE(desc && counts && bow_words && bow_vals && nbow, "bowx_transform_batch: NULL pointer");
    const bool fv = fv_nodes || fv_offsets || fv_feats || nfv;
    ORBX_REQUIRE(!fv || (fv_nodes && fv_offsets && fv_feats && nfv), "bowx_transform_batch: the feature vector needs all four of its arrays");
    for (int i = 0; i < nframes; i++)
        ORBX_REQUIRE(counts[i] >= 0 && counts[i] <= cap, "bowx_transform_batch: counts[%d] = %d outside [0, cap = %d]", i, counts[i], cap);
    ORBX_CUDA(cudaSetDevice(h->device));
    const size_t total = (size_t)nframes * cap, nf = (size_t)nframes;
    const size_t need = align_up(total * 32, 256) + align_up(nf * 4, 256) * 3 + align_up(total * 4, 256) * 3 + align_up(total * 8, 256) +
                        align_up(nf * ((size_t)cap + 1) * 4, 256);
    rc = bow_grow(&h->d_buf, &h->buf_bytes, need);
    if (rc) return rc;
    uint8_t* p = h->d_buf;
    uint8_t* d_desc = carve(p, total * 32);
    int32_t* d_counts = (int32_t*)carve(p, nf * 4);
    int32_t* d_nbow = (int32_t*)carve(p, nf * 4);
    int32_t* d_nfv = (int32_t*)carve(p, nf * 4);
    uint32_t* d_words = (uint32_t*)carve(p, total * 4);
    uint32_t* d_nodes = (uint32_t*)carve(p, total * 4);
    uint32_t* d_feats = (uint32_t*)carve(p, total * 4);
    double* d_vals = (double*)carve(p, total * 8);
    int32_t* d_offs = (int32_t*)carve(p, nf * ((size_t)cap + 1) * 4);
    ORBX_CUDA(cudaMemcpyAsync(d_desc, desc, total * 32, cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(d_counts, counts, nf * 4, cudaMemcpyHostToDevice, h->stream));
    rc = bowx_transform_batch_dev(h, d_desc, d_counts, nframes, cap, levelsup, d_words, d_vals, d_nbow, fv ? d_nodes : nullptr,
                                  fv ? d_offs : nullptr, fv ? d_feats : nullptr, fv ? d_nfv : nullptr);
    if (rc) return rc;
    ORBX_CUDA(cudaMemcpyAsync(bow_words, d_words, total * 4, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(bow_vals, d_vals, total * 8, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(nbow, d_nbow, nf * 4, cudaMemcpyDeviceToHost, h->stream));
    if (fv) {
        ORBX_CUDA(cudaMemcpyAsync(fv_nodes, d_nodes, total * 4, cudaMemcpyDeviceToHost, h->stream));
        ORBX_CUDA(cudaMemcpyAsync(fv_offsets, d_offs, nf * ((size_t)cap + 1) * 4, cudaMemcpyDeviceToHost, h->stream));
        ORBX_CUDA(cudaMemcpyAsync(fv_feats, d_feats, total * 4, cudaMemcpyDeviceToHost, h->stream));
        ORBX_CUDA(cudaMemcpyAsync(nfv, d_nfv, nf * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int bowx_score_batch_dev(bowx_handle h, const uint32_t* d_qwords, const double* d_qvals, int nq, const int64_t* d_db_start,
                                    const int32_t* d_db_count, const uint32_t* d_db_words, const double* d_db_vals, int nentries,
                                    double* d_scores)
{
    ORBX_REQUIRE(h != nullptr, "bowx_score_batch_dev: NULL handle");
    ORBX_REQUIRE(h->nnodes > 0, "bowx_score_batch_dev: no vocabulary (its scoring type selects the score)");
    ORBX_REQUIRE(nq >= 0 && nentries >= 0, "bowx_score_batch_dev: nq %d / nentries %d", nq, nentries);
    if (nentries == 0) return ORBX_OK;
    ORBX_REQUIRE((nq == 0 || (d_qwords && d_qvals)) && d_db_start && d_db_count && d_db_words && d_db_vals && d_scores, "bowx_score_batch_dev: NULL pointer");
    ORBX_REQUIRE(nq <= 8192, "bowx_score_batch_dev: a query of %d words (at most 8192)", nq);
    const int vals_in_smem = nq <= 3584;         // tables 18 KB + words + values within the 64 KB the kernels opt in to
    const size_t smem = ((size_t)BOW_TABLE_WORDS + (size_t)std::max(nq, 1) * (vals_in_smem ? 3 : 1)) * sizeof(uint32_t);
    ORBX_CUDA(cudaSetDevice(h->device));
    // a CTA copies the query's tables once and its 8 warps take stored vectors from a counter until none is left
    const unsigned blocks = (unsigned)std::min((nentries + BOW_THREADS / 32 - 1) / (BOW_THREADS / 32), h->sm_count * 4);
    int word_bits = 1;
    while (word_bits < 31 && (1u << word_bits) < (unsigned)std::max(h->nwords, 2)) word_bits++;
    const int word_shift = std::max(0, word_bits - 10);          // BOW_BUCKETS ranges cover the word ids
    k_bow_query_tables<<<1, 1024, 0, h->stream>>>(d_qwords, nq, word_shift, h->d_tables, h->d_next);
    ORBX_CUDA(cudaGetLastError());
#define BOW_SCORE(S) k_bow_score<S><<<blocks, BOW_THREADS, smem, h->stream>>>(d_qwords, d_qvals, nq, h->d_tables, d_db_start, d_db_count, d_db_words, d_db_vals, nentries, word_shift, vals_in_smem, h->d_next, d_scores)
    switch (h->scoring) {
    case BOWX_L1_NORM: BOW_SCORE(S_L1); break;
    case BOWX_L2_NORM: BOW_SCORE(S_L2); break;
    case BOWX_CHI_SQUARE: BOW_SCORE(S_CHI); break;
    case BOWX_KL: BOW_SCORE(S_KL); break;
    case BOWX_BHATTACHARYYA: BOW_SCORE(S_BHAT); break;
    default: BOW_SCORE(S_DOT); break;
    }
#undef BOW_SCORE
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int bowx_score_batch(bowx_handle h, const uint32_t* qwords, const double* qvals, int nq, const int64_t* db_start,
                                const int32_t* db_count, const uint32_t* db_words, const double* db_vals, int64_t db_len, int nentries,
                                double* scores)
{
    ORBX_REQUIRE(h != nullptr, "bowx_score_batch: NULL handle");
    ORBX_REQUIRE(nq >= 0 && nentries >= 0 && db_len >= 0, "bowx_score_batch: nq %d / nentries %d / db_len %lld", nq, nentries, (long long)db_len);
    if (nentries == 0) return ORBX_OK;
    ORBX_REQUIRE((nq == 0 || (qwords && qvals)) && db_start && db_count && (db_len == 0 || (db_words && db_vals)) && scores, "bowx_score_batch: NULL pointer");
    for (int e = 0; e < nentries; e++)
        ORBX_REQUIRE(db_count[e] >= 0 && db_start[e] >= 0 && db_start[e] + db_count[e] <= db_len,
                     "bowx_score_batch: entry %d [%lld, +%d) outside the %lld stored words", e, (long long)db_start[e], db_count[e], (long long)db_len);
    ORBX_CUDA(cudaSetDevice(h->device));
    const size_t ne = (size_t)nentries, q = (size_t)std::max(nq, 1), dl = (size_t)std::max<int64_t>(db_len, 1);
    const size_t need = align_up(q * 4, 256) + align_up(q * 8, 256) + align_up(ne * 8, 256) * 2 + align_up(ne * 4, 256) + align_up(dl * 4, 256) +
                        align_up(dl * 8, 256);
    int rc = bow_grow(&h->d_buf, &h->buf_bytes, need);
    if (rc) return rc;
    uint8_t* p = h->d_buf;
    uint32_t* d_qw = (uint32_t*)carve(p, q * 4);
    double* d_qv = (double*)carve(p, q * 8);
    int64_t* d_start = (int64_t*)carve(p, ne * 8);
    double* d_scores = (double*)carve(p, ne * 8);
    int32_t* d_cnt = (int32_t*)carve(p, ne * 4);
    uint32_t* d_w = (uint32_t*)carve(p, dl * 4);
    double* d_v = (double*)carve(p, dl * 8);
    if (nq) {
        ORBX_CUDA(cudaMemcpyAsync(d_qw, qwords, (size_t)nq * 4, cudaMemcpyHostToDevice, h->stream));
        ORBX_CUDA(cudaMemcpyAsync(d_qv, qvals, (size_t)nq * 8, cudaMemcpyHostToDevice, h->stream));
    }
    ORBX_CUDA(cudaMemcpyAsync(d_start, db_start, ne * 8, cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(d_cnt, db_count, ne * 4, cudaMemcpyHostToDevice, h->stream));
    if (db_len) {
        ORBX_CUDA(cudaMemcpyAsync(d_w, db_words, (size_t)db_len * 4, cudaMemcpyHostToDevice, h->stream));
        ORBX_CUDA(cudaMemcpyAsync(d_v, db_vals, (size_t)db_len * 8, cudaMemcpyHostToDevice, h->stream));
    }
    rc = bowx_score_batch_dev(h, d_qw, d_qv, nq, d_start, d_cnt, d_w, d_v, nentries, d_scores);
    if (rc) return rc;
    ORBX_CUDA(cudaMemcpyAsync(scores, d_scores, ne * 8, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int bowx_score(bowx_handle h, const uint32_t* words1, const double* vals1, int n1, const uint32_t* words2, const double* vals2,
                          int n2, double* score)
{
    ORBX_REQUIRE(score != nullptr && n2 >= 0, "bowx_score: bad argument");
    const int64_t start = 0;
    const int32_t count = n2;
    return bowx_score_batch(h, words1, vals1, n1, &start, &count, words2, vals2, n2, 1, score);
}
