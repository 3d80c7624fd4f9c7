// triangulate.cu -- the consumers of the match lists in CameraPoseEstimator, batched on the device (sm_100a):
//
//   K8  k_triangulate     TriangulateSinglePointFromTwoView for every correspondence of every problem, one thread each
//                         (reference src/CameraPoseEstimator.cpp:86-132; called per match at :506 and, four times over the
//                         inliers, by the bootstrap's hypothesis test :334-349 through TriangulateMultiplePointsFromTwoView
//                         :134-152): P = K [R|t] of both views, the 4x4 system of :94-111, its null direction (the right
//                         singular vector of the smallest singular value -- solveHLS, src/CommonMath.cpp:17-22), X / X[3],
//                         and the front-of-both-cameras flag :119-127.
//   K9  k_associate       the association loop of pnpPoseEstimation (:402-455): which current feature inherits which map
//                         point from which predecessor, first come first served, and the list of associations in the order
//                         the reference discovers them.
//   K10 k_select_new      the new-map-point loop (:488-512): the greedy, order-dependent choice of the matches that are
//                         triangulated, as a parallel fixed point.
//
// The singular vector is computed in double by the one-sided Jacobi method (Hestenes) on the columns of A, entirely in
// registers: the 4x4 problem is too small for anything else, and unlike an eigen-decomposition of A^T A it does not square
// the condition number (pixel coordinates put sigma_1 / sigma_3 near 1e3..1e4).  cv::SVD is the same method.
#include <float.h>
#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace orbx {
namespace {

constexpr int TR_THREADS = 128;

struct PtsF32 {   // [prob][cap][2] float, the layout of fmx_fundamental_batch's inputs
    const float* p1; const float* p2; int cap;
    __device__ bool get(int prob, int i, double& x1, double& y1, double& x2, double& y2) const
    {
        const float2 a = reinterpret_cast<const float2*>(p1)[(size_t)prob * cap + i];
        const float2 b = reinterpret_cast<const float2*>(p2)[(size_t)prob * cap + i];
        x1 = a.x; y1 = a.y; x2 = b.x; y2 = b.y;
        return true;
    }
};
struct PtsF64 {   // [prob][cap][2] double (vector<Point2d>, the reference's own argument type)
    const double* p1; const double* p2; int cap;
    __device__ bool get(int prob, int i, double& x1, double& y1, double& x2, double& y2) const
    {
        const double2 a = reinterpret_cast<const double2*>(p1)[(size_t)prob * cap + i];
        const double2 b = reinterpret_cast<const double2*>(p2)[(size_t)prob * cap + i];
        x1 = a.x; y1 = a.y; x2 = b.x; y2 = b.y;
        return true;
    }
};
// pair (f, j) of the steady-state loop: match i of list f*back + j-1 joins keypoint train_idx of frame f-j (view 1, as at
// :504-506) and keypoint query_idx of frame f (view 2).  Predecessors before the batch come from the keypoint history.
struct PtsBack {
    const orbx_keypoint* kps; const orbx_keypoint* hist; const orbx_dmatch* good; int cap, back, nhist;
    __device__ bool get(int prob, int i, double& x1, double& y1, double& x2, double& y2) const
    {
        const int f = prob / back, j = prob - f * back + 1, src = f - j;
        const orbx_keypoint* pre = src >= 0 ? kps + (size_t)src * cap : (-src - 1 < nhist ? hist + (size_t)(-src - 1) * cap : nullptr);
        if (!pre) return false;
        const orbx_dmatch m = good[(size_t)prob * cap + i];
        const orbx_keypoint a = pre[m.train_idx], b = kps[(size_t)f * cap + m.query_idx];
        x1 = a.x; y1 = a.y; x2 = b.x; y2 = b.y;
        return true;
    }
};

// One Jacobi rotation of rows i and j of At (and of Vt); w = squared row norms.  Same formulae as the textbook method.
#define TR_ROTATE(i, j)                                                                                  \
    do {                                                                                                 \
        double p = At[i][0] * At[j][0] + At[i][1] * At[j][1] + At[i][2] * At[j][2] + At[i][3] * At[j][3]; \
        const double a = w[i], b = w[j];                                                                 \
        if (fabs(p) > eps * sqrt(a * b)) {                                                               \
            p *= 2;                                                                                      \
            const double beta = a - b, gamma = hypot(p, beta);                                           \
            double c, s;                                                                                 \
            if (beta < 0) {                                                                              \
                const double delta = (gamma - beta) * 0.5;                                               \
                s = sqrt(delta / gamma);                                                                 \
                c = p / (gamma * s * 2);                                                                 \
            } else {                                                                                     \
                c = sqrt((gamma + beta) / (gamma * 2));                                                  \
                s = p / (gamma * c * 2);                                                                 \
            }                                                                                            \
            double na = 0, nb = 0;                                                                       \
            _Pragma("unroll") for (int k = 0; k < 4; k++) {                                              \
                const double t0 = c * At[i][k] + s * At[j][k], t1 = c * At[j][k] - s * At[i][k];         \
                At[i][k] = t0; At[j][k] = t1;                                                            \
                na += t0 * t0; nb += t1 * t1;                                                            \
                const double v0 = c * Vt[i][k] + s * Vt[j][k], v1 = c * Vt[j][k] - s * Vt[i][k];         \
                Vt[i][k] = v0; Vt[j][k] = v1;                                                            \
            }                                                                                            \
            w[i] = na; w[j] = nb;                                                                        \
            changed = true;                                                                              \
        }                                                                                                \
    } while (0)

// grid = (ceil(cap / 128), nprob * nhyp).  X is [nprob * nhyp][cap][3], front [nprob * nhyp][cap], nfront [nprob * nhyp]
// (zeroed by the host before the launch).  Entries that are not selected (or lie beyond the problem's count) get X = 0,
// front = 0, so the arrays are fully defined.
template <class Pts, typename CountT>
__global__ void __launch_bounds__(TR_THREADS)
k_triangulate(const Pts pts, const CountT* __restrict__ counts, const uint8_t* __restrict__ select, int cap, const trx_cameras* __restrict__ cams,
              int nhyp, double* __restrict__ X, uint8_t* __restrict__ front, int32_t* __restrict__ nfront)
{
    __shared__ double sP[2][12];     // K1 Rt1, K2 Rt2
    __shared__ double sZ[2][4];      // third rows of Rt1, Rt2 (depth in each camera)
    __shared__ int s_count;
    const int ph = blockIdx.y, prob = ph / nhyp;
    const int n = min((int)counts[prob], cap);
    const int i = blockIdx.x * TR_THREADS + threadIdx.x;
    const trx_cameras& cam = cams[ph];
    if (threadIdx.x < 24) {
        const int v = threadIdx.x / 12, e = threadIdx.x % 12, r = e / 4, c = e % 4;
        const double* K = v ? cam.K2 : cam.K1;
        const double* Rt = v ? cam.Rt2 : cam.Rt1;
        sP[v][e] = K[3 * r] * Rt[c] + K[3 * r + 1] * Rt[4 + c] + K[3 * r + 2] * Rt[8 + c];
    } else if (threadIdx.x < 32) {
        const int v = (threadIdx.x - 24) / 4, c = (threadIdx.x - 24) % 4;
        sZ[v][c] = (v ? cam.Rt2 : cam.Rt1)[8 + c];
    }
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    if (i >= cap) return;

    double x1, y1, x2, y2;
    bool ok = i < n && (select == nullptr || select[(size_t)prob * cap + i] != 0);
    if (ok) ok = pts.get(prob, i, x1, y1, x2, y2);
    double Xo[3] = { 0., 0., 0. };
    int in_front = 0;
    if (ok) {
        // At[c][r] = A[r][c]: the columns of A are rotated against each other
        double At[4][4], Vt[4][4], w[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            At[c][0] = sP[0][c] - sP[0][8 + c] * x1;
            At[c][1] = sP[0][4 + c] - sP[0][8 + c] * y1;
            At[c][2] = sP[1][c] - sP[1][8 + c] * x2;
            At[c][3] = sP[1][4 + c] - sP[1][8 + c] * y2;
            w[c] = At[c][0] * At[c][0] + At[c][1] * At[c][1] + At[c][2] * At[c][2] + At[c][3] * At[c][3];
#pragma unroll
            for (int k = 0; k < 4; k++) Vt[c][k] = c == k ? 1. : 0.;
        }
        const double eps = DBL_EPSILON * 10;
        for (int iter = 0; iter < 30; iter++) {
            bool changed = false;
            TR_ROTATE(0, 1); TR_ROTATE(0, 2); TR_ROTATE(0, 3);
            TR_ROTATE(1, 2); TR_ROTATE(1, 3); TR_ROTATE(2, 3);
            if (!changed) break;
        }
        // the column with the smallest norm spans the null direction; on equal norms the later one (a descending sort that
        // only swaps on strict '<' leaves it last)
        int m = 0;
        double wm = At[0][0] * At[0][0] + At[0][1] * At[0][1] + At[0][2] * At[0][2] + At[0][3] * At[0][3];
#pragma unroll
        for (int c = 1; c < 4; c++) {
            const double wc = At[c][0] * At[c][0] + At[c][1] * At[c][1] + At[c][2] * At[c][2] + At[c][3] * At[c][3];
            if (wc <= wm) { wm = wc; m = c; }
        }
        double v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = m == 0 ? Vt[0][k] : m == 1 ? Vt[1][k] : m == 2 ? Vt[2][k] : Vt[3][k];
        const double Xh[4] = { v[0] / v[3], v[1] / v[3], v[2] / v[3], v[3] / v[3] };
        Xo[0] = Xh[0]; Xo[1] = Xh[1]; Xo[2] = Xh[2];
        double z1 = 0, z2 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { z1 += sZ[0][k] * Xh[k]; z2 += sZ[1][k] * Xh[k]; }
        in_front = z1 > 0 && z2 > 0;
    }
    const size_t o = (size_t)ph * cap + i;
    X[3 * o] = Xo[0]; X[3 * o + 1] = Xo[1]; X[3 * o + 2] = Xo[2];
    front[o] = (uint8_t)in_front;
    const unsigned int bal = __ballot_sync(__activemask(), in_front);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&s_count, __popc(bal));
    __syncthreads();
    if (threadIdx.x == 0 && s_count) atomicAdd(&nfront[ph], s_count);
}

// :342-347: `if (maxCount < count)` walking the hypotheses in order = the first maximum.
__global__ void k_tri_best(const int32_t* __restrict__ nfront, int nprob, int nhyp, int32_t* __restrict__ best)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nprob) return;
    int b = -1, mx = -1;
    for (int h = 0; h < nhyp; h++) {
        const int c = nfront[p * nhyp + h];
        if (mx < c) { mx = c; b = h; }
    }
    best[p] = b;
}

// ---------------------------------------------------------------------------------------------- association (K9)
// One CTA per problem (= one current frame and its `back` match lists).  List l of problem p: matches good[(p*back + l)*cap ..],
// ngood[p*back + l] of them, optional RANSAC status bytes (only set entries take part -- the FILTERING_WITH_F block :412-424
// drops the others before the loop), premap[(p*back + l)*cap + t] = map point of the predecessor's feature t or -1.
// Order of the reference's walk: list, then position.  key = l * cap + pos; the earliest match that offers a map point to a
// current feature wins it (`!matched[cur]`, :436), found with an atomic minimum; the association list is then compacted in
// key order with a block-wide scan, which is the order of mapPoints / imagePoints at :447-449.
constexpr int AS_THREADS = 512;
constexpr uint32_t AS_NONE = 0xFFFFFFFFu;

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int before = 0, tot = 0;
    for (int k = 0; k < nw; k++) {
        const int s = s_warp[k];
        before += k < wid ? s : 0;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return before + incl - v;
}

__global__ void __launch_bounds__(AS_THREADS)
k_associate(const orbx_dmatch* __restrict__ good, const long long* __restrict__ ngood, const uint8_t* __restrict__ status,
            const int32_t* __restrict__ premap, const int32_t* __restrict__ ncur, int back, int cap, int32_t* __restrict__ cur_map,
            int32_t* __restrict__ assoc_q, int32_t* __restrict__ assoc_mp, int32_t* __restrict__ nassoc)
{
    __shared__ int s_warp[AS_THREADS / 32];
    const int p = blockIdx.x;
    uint32_t* key = reinterpret_cast<uint32_t*>(cur_map) + (size_t)p * cap;       // scratch first, the result at the end
    const int nc = min(ncur[p], cap);
    for (int i = threadIdx.x; i < cap; i += blockDim.x) key[i] = AS_NONE;
    __syncthreads();
    for (int l = 0; l < back; l++) {
        const size_t base = ((size_t)p * back + l) * cap;
        const int n = (int)min((long long)cap, ngood[(size_t)p * back + l]);
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            if (status && !status[base + j]) continue;
            const orbx_dmatch m = good[base + j];
            if ((unsigned)m.query_idx >= (unsigned)nc || (unsigned)m.train_idx >= (unsigned)cap) continue;
            if (premap[base + m.train_idx] != -1) atomicMin(&key[m.query_idx], (uint32_t)(l * cap + j));
        }
    }
    __syncthreads();
    int written = 0;
    for (int l = 0; l < back; l++) {
        const size_t base = ((size_t)p * back + l) * cap;
        const int n = (int)min((long long)cap, ngood[(size_t)p * back + l]);
        for (int c = 0; c < n; c += blockDim.x) {
            const int j = c + threadIdx.x;
            int take = 0, q = 0, mp = -1;
            if (j < n && !(status && !status[base + j])) {
                const orbx_dmatch m = good[base + j];
                if ((unsigned)m.query_idx < (unsigned)nc && (unsigned)m.train_idx < (unsigned)cap) {
                    q = m.query_idx;
                    mp = premap[base + m.train_idx];
                    take = mp != -1 && key[q] == (uint32_t)(l * cap + j);
                }
            }
            int total;
            const int pos = block_exclusive_scan(take, s_warp, total);
            if (take) {
                assoc_q[(size_t)p * cap + written + pos] = q;
                assoc_mp[(size_t)p * cap + written + pos] = mp;
            }
            written += total;
        }
    }
    __syncthreads();
    // key -> map point of the winning match
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        const uint32_t k = key[i];
        int32_t mp = -1;
        if (k != AS_NONE) {
            const int l = (int)(k / (uint32_t)cap), j = (int)(k - (uint32_t)l * cap);
            const size_t base = ((size_t)p * back + l) * cap;
            mp = premap[base + good[base + j].train_idx];
        }
        cur_map[(size_t)p * cap + i] = mp;
    }
    if (threadIdx.x == 0) nassoc[p] = written;
}

// ---------------------------------------------------------------------------------------------- new map points (K10)
// :488-512 walks the lists in the same order and takes a match iff neither of its two features has a map point yet; taking it
// gives both features the new point (registerNewMapPoint :235-243), so it blocks every later match that shares its current
// feature (any list) or its predecessor feature (same list).  That is a greedy matching in a fixed order.  In parallel: a
// candidate that is the earliest live candidate on BOTH of its features is certainly taken (everything before it on those
// features is dead); a candidate sharing a feature with a taken one is dead; repeat until nothing is undecided.  Each round
// decides at least the earliest undecided candidate, and conflicts are rare, so two or three rounds are typical.
//   state per match (the `accept` output doubles as it): 0 = dead / not a candidate, 1 = taken, 2 = undecided
// eq [p][cap] and et [p][back][cap] hold the earliest live key per current / predecessor feature (workspace).
__global__ void __launch_bounds__(AS_THREADS)
k_select_new(const orbx_dmatch* __restrict__ good, const long long* __restrict__ ngood, const uint8_t* __restrict__ status,
             int32_t* __restrict__ premap, int32_t* __restrict__ cur_map, const int32_t* __restrict__ ncur, const int32_t* __restrict__ next_id,
             int back, int cap, uint8_t* __restrict__ accept, uint32_t* __restrict__ eq_all, uint32_t* __restrict__ et_all,
             int32_t* __restrict__ nnew)
{
    __shared__ int s_warp[AS_THREADS / 32];
    __shared__ int s_undecided;
    const int p = blockIdx.x;
    const int nc = min(ncur[p], cap);
    uint32_t* eq = eq_all + (size_t)p * cap;
    uint32_t* et = et_all + (size_t)p * back * cap;
    const size_t pbase = (size_t)p * back * cap;
    // candidates: both features free at the start
    for (int l = 0; l < back; l++) {
        const size_t base = pbase + (size_t)l * cap;
        const int n = (int)min((long long)cap, ngood[(size_t)p * back + l]);
        for (int j = threadIdx.x; j < cap; j += blockDim.x) {
            uint8_t st = 0;
            if (j < n && !(status && !status[base + j])) {
                const orbx_dmatch m = good[base + j];
                if ((unsigned)m.query_idx < (unsigned)nc && (unsigned)m.train_idx < (unsigned)cap &&
                    premap[base + m.train_idx] == -1 && cur_map[(size_t)p * cap + m.query_idx] == -1)
                    st = 2;
            }
            accept[base + j] = st;
        }
    }
    __syncthreads();
    for (int round = 0; round < back * cap + 1; round++) {
        for (int i = threadIdx.x; i < cap; i += blockDim.x) eq[i] = AS_NONE;
        for (int i = threadIdx.x; i < back * cap; i += blockDim.x) et[i] = AS_NONE;
        if (threadIdx.x == 0) s_undecided = 0;
        __syncthreads();
        // earliest live (taken or undecided) candidate per feature
        for (int l = 0; l < back; l++) {
            const size_t base = pbase + (size_t)l * cap;
            const int n = (int)min((long long)cap, ngood[(size_t)p * back + l]);
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                if (!accept[base + j]) continue;
                const orbx_dmatch m = good[base + j];
                atomicMin(&eq[m.query_idx], (uint32_t)(l * cap + j));
                atomicMin(&et[(size_t)l * cap + m.train_idx], (uint32_t)(l * cap + j));
            }
        }
        __syncthreads();
        // earliest on both features -> taken
        for (int l = 0; l < back; l++) {
            const size_t base = pbase + (size_t)l * cap;
            const int n = (int)min((long long)cap, ngood[(size_t)p * back + l]);
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                if (accept[base + j] != 2) continue;
                const orbx_dmatch m = good[base + j];
                const uint32_t k = (uint32_t)(l * cap + j);
                if (eq[m.query_idx] == k && et[(size_t)l * cap + m.train_idx] == k) accept[base + j] = 1;
            }
        }
        __syncthreads();
        // sharing a feature with a taken candidate -> dead (a taken candidate is always the earliest live one on its features)
        for (int l = 0; l < back; l++) {
            const size_t base = pbase + (size_t)l * cap;
            const int n = (int)min((long long)cap, ngood[(size_t)p * back + l]);
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                if (accept[base + j] != 2) continue;
                const orbx_dmatch m = good[base + j];
                const uint32_t kq = eq[m.query_idx], kt = et[(size_t)l * cap + m.train_idx];
                const int lq = (int)(kq / (uint32_t)cap), lt = (int)(kt / (uint32_t)cap);
                const bool dead = accept[pbase + (size_t)lq * cap + (kq - (uint32_t)lq * cap)] == 1 ||
                                  accept[pbase + (size_t)lt * cap + (kt - (uint32_t)lt * cap)] == 1;
                if (dead) accept[base + j] = 0;
                else s_undecided = 1;
            }
        }
        __syncthreads();
        const int more = s_undecided;
        __syncthreads();
        if (!more) break;
    }
    // ids in the order of the walk, and the two index tables updated as registerNewMapPoint does
    const int id0 = next_id ? next_id[p] : 0;
    int written = 0;
    for (int l = 0; l < back; l++) {
        const size_t base = pbase + (size_t)l * cap;
        const int n = (int)min((long long)cap, ngood[(size_t)p * back + l]);
        for (int c = 0; c < n; c += blockDim.x) {
            const int j = c + threadIdx.x;
            const int take = j < n && accept[base + j] == 1;
            int total;
            const int pos = block_exclusive_scan(take, s_warp, total);
            if (take) {
                const orbx_dmatch m = good[base + j];
                premap[base + m.train_idx] = id0 + written + pos;
                cur_map[(size_t)p * cap + m.query_idx] = id0 + written + pos;
            }
            written += total;
        }
    }
    if (threadIdx.x == 0) nnew[p] = written;
}

}  // namespace
}  // namespace orbx

// ------------------------------------------------------------------------------------------------ host side
using namespace orbx;

struct trx_context {
    int device;
    cudaStream_t own_stream, stream;
    // staging for the host-buffer entry points, and the scratch of k_select_new (grown on demand)
    uint8_t* d_buf; size_t buf_bytes;
    uint32_t* d_scratch; size_t scratch_bytes;
    int32_t* h_small;      // pinned: counts read back by the host forms
};

static int trx_grow(void** p, size_t* have, size_t want)
{
    if (*p && *have >= want) return ORBX_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *have = 0; }
    const size_t bytes = align_up(want + want / 4, 256);
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) { *p = nullptr; set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return ORBX_E_ALLOC; }
    *have = bytes;
    return ORBX_OK;
}

extern "C" int trx_create(trx_handle* out, int device)
{
    ORBX_REQUIRE(out != nullptr, "trx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { set_error("trx_create: no CUDA device (%s); liborbx has no CPU fallback", cudaGetErrorString(e)); return ORBX_E_CUDA; }
    ORBX_REQUIRE(device >= 0 && device < ndev, "trx_create: device %d out of range [0,%d)", device, ndev);
    ORBX_CUDA(cudaSetDevice(device));
    trx_context* h = new trx_context();
    memset(h, 0, sizeof(*h));
    h->device = device;
    ORBX_CUDA_OR(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking), trx_destroy(h));
    h->stream = h->own_stream;
    ORBX_CUDA_OR(cudaMallocHost((void**)&h->h_small, 4096), trx_destroy(h));
    *out = h;
    return ORBX_OK;
}

extern "C" int trx_destroy(trx_handle h)
{
    if (!h) return ORBX_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_buf); cudaFree(h->d_scratch);
    if (h->h_small) cudaFreeHost(h->h_small);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return ORBX_OK;
}

extern "C" int trx_set_stream(trx_handle h, void* cuda_stream)
{
    ORBX_REQUIRE(h != nullptr, "trx_set_stream: NULL handle");
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return ORBX_OK;
}

extern "C" int trx_get_stream(trx_handle h, void** cuda_stream)
{
    ORBX_REQUIRE(h != nullptr && cuda_stream != nullptr, "trx_get_stream: NULL argument");
    *cuda_stream = (void*)h->stream;
    return ORBX_OK;
}

extern "C" int trx_synchronize(trx_handle h)
{
    ORBX_REQUIRE(h != nullptr, "trx_synchronize: NULL handle");
    ORBX_CUDA(cudaSetDevice(h->device));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

template <class Pts, typename CountT>
static int launch_triangulate(trx_handle h, const Pts& pts, const CountT* d_counts, const uint8_t* d_select, int nprob, int cap,
                              const trx_cameras* d_cams, int nhyp, double* d_X, uint8_t* d_front, int32_t* d_nfront, int32_t* d_best)
{
    ORBX_CUDA(cudaMemsetAsync(d_nfront, 0, (size_t)nprob * nhyp * sizeof(int32_t), h->stream));
    const dim3 grid((unsigned)div_up(cap, TR_THREADS), (unsigned)(nprob * nhyp));
    k_triangulate<Pts, CountT><<<grid, TR_THREADS, 0, h->stream>>>(pts, d_counts, d_select, cap, d_cams, nhyp, d_X, d_front, d_nfront);
    ORBX_CUDA(cudaGetLastError());
    if (d_best) {
        k_tri_best<<<div_up(nprob, 128), 128, 0, h->stream>>>(d_nfront, nprob, nhyp, d_best);
        ORBX_CUDA(cudaGetLastError());
    }
    return ORBX_OK;
}

static int tri_checks(trx_handle h, int nprob, int cap, int nhyp, const void* a, const void* b, const void* c, const void* d, const char* fn)
{
    ORBX_REQUIRE(h != nullptr, "%s: NULL handle", fn);
    ORBX_REQUIRE(nprob >= 0 && cap >= 1 && nhyp >= 1 && nhyp <= 64 && (long long)nprob * nhyp <= 65535, "%s: bad sizes (nprob %d, cap %d, nhyp %d)", fn, nprob, cap, nhyp);
    ORBX_REQUIRE(nprob == 0 || (a && b && c && d), "%s: NULL pointer", fn);
    ORBX_CUDA(cudaSetDevice(h->device));
    return ORBX_OK;
}

extern "C" int trx_triangulate_batch_dev(trx_handle h, const float* d_pts1, const float* d_pts2, const int32_t* d_counts, const uint8_t* d_select,
                                         int nprob, int cap, const trx_cameras* d_cams, int nhyp, double* d_X, uint8_t* d_front,
                                         int32_t* d_nfront, int32_t* d_best)
{
    int rc = tri_checks(h, nprob, cap, nhyp, d_pts1, d_pts2, d_counts, d_cams, "trx_triangulate_batch_dev");
    if (rc || nprob == 0) return rc;
    ORBX_REQUIRE(d_X && d_front && d_nfront, "trx_triangulate_batch_dev: NULL output");
    if ((((uintptr_t)d_pts1) | ((uintptr_t)d_pts2) | ((uintptr_t)d_cams) | ((uintptr_t)d_X)) & 7) { set_error("trx_triangulate_batch_dev: device pointers must be 8-byte aligned"); return ORBX_E_ALIGN; }
    const PtsF32 pts = { d_pts1, d_pts2, cap };
    return launch_triangulate(h, pts, d_counts, d_select, nprob, cap, d_cams, nhyp, d_X, d_front, d_nfront, d_best);
}

extern "C" int trx_triangulate_back_dev(trx_handle h, const orbx_keypoint* d_kps, int nframes, int cap, int back, const orbx_keypoint* d_hist_kps,
                                        int nhist, const orbx_dmatch* d_good, const int64_t* d_ngood, const uint8_t* d_select,
                                        const trx_cameras* d_cams, double* d_X, uint8_t* d_front, int32_t* d_nfront)
{
    ORBX_REQUIRE(back >= 1 && back <= 64 && nframes >= 0 && nhist >= 0, "trx_triangulate_back_dev: bad sizes");
    int rc = tri_checks(h, nframes * back, cap, 1, d_kps, d_good, d_ngood, d_cams, "trx_triangulate_back_dev");
    if (rc || nframes == 0) return rc;
    ORBX_REQUIRE(d_X && d_front && d_nfront && (nhist == 0 || d_hist_kps), "trx_triangulate_back_dev: NULL pointer");
    const PtsBack pts = { d_kps, d_hist_kps, d_good, cap, back, nhist };
    return launch_triangulate(h, pts, reinterpret_cast<const long long*>(d_ngood), d_select, nframes * back, cap, d_cams, 1, d_X, d_front,
                              d_nfront, nullptr);
}

// host buffers: one staging allocation laid out [pts1 | pts2 | counts | select | cams | X | front | nfront | best]
template <typename PtT>
static int triangulate_host(trx_handle h, const PtT* pts1, const PtT* pts2, const int32_t* counts, const uint8_t* select, int nprob, int cap,
                            const trx_cameras* cams, int nhyp, double* X, uint8_t* front, int32_t* nfront, int32_t* best, const char* fn)
{
    int rc = tri_checks(h, nprob, cap, nhyp, pts1, pts2, counts, cams, fn);
    if (rc || nprob == 0) return rc;
    ORBX_REQUIRE(X && nfront, "%s: NULL output", fn);
    const size_t np = (size_t)nprob, nph = np * nhyp;
    const size_t o_p1 = 0, o_p2 = o_p1 + align_up(np * cap * 2 * sizeof(PtT), 256), o_cnt = o_p2 + align_up(np * cap * 2 * sizeof(PtT), 256);
    const size_t o_sel = o_cnt + align_up(np * 4, 256), o_cam = o_sel + align_up(np * cap, 256), o_X = o_cam + align_up(nph * sizeof(trx_cameras), 256);
    const size_t o_fr = o_X + align_up(nph * cap * 24, 256), o_nf = o_fr + align_up(nph * cap, 256), o_best = o_nf + align_up(nph * 4, 256);
    const size_t total = o_best + align_up(np * 4, 256);
    rc = trx_grow((void**)&h->d_buf, &h->buf_bytes, total);
    if (rc) return rc;
    uint8_t* d = h->d_buf;
    cudaStream_t s = h->stream;
    ORBX_CUDA(cudaMemcpyAsync(d + o_p1, pts1, np * cap * 2 * sizeof(PtT), cudaMemcpyHostToDevice, s));
    ORBX_CUDA(cudaMemcpyAsync(d + o_p2, pts2, np * cap * 2 * sizeof(PtT), cudaMemcpyHostToDevice, s));
    ORBX_CUDA(cudaMemcpyAsync(d + o_cnt, counts, np * 4, cudaMemcpyHostToDevice, s));
    if (select) ORBX_CUDA(cudaMemcpyAsync(d + o_sel, select, np * cap, cudaMemcpyHostToDevice, s));
    ORBX_CUDA(cudaMemcpyAsync(d + o_cam, cams, nph * sizeof(trx_cameras), cudaMemcpyHostToDevice, s));
    typename std::conditional<sizeof(PtT) == 8, PtsF64, PtsF32>::type pts;
    pts.p1 = (const PtT*)(d + o_p1); pts.p2 = (const PtT*)(d + o_p2); pts.cap = cap;
    rc = launch_triangulate(h, pts, (const int32_t*)(d + o_cnt), select ? d + o_sel : nullptr, nprob, cap, (const trx_cameras*)(d + o_cam), nhyp,
                            (double*)(d + o_X), d + o_fr, (int32_t*)(d + o_nf), best ? (int32_t*)(d + o_best) : nullptr);
    if (rc) return rc;
    ORBX_CUDA(cudaMemcpyAsync(X, d + o_X, nph * cap * 24, cudaMemcpyDeviceToHost, s));
    if (front) ORBX_CUDA(cudaMemcpyAsync(front, d + o_fr, nph * cap, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(nfront, d + o_nf, nph * 4, cudaMemcpyDeviceToHost, s));
    if (best) ORBX_CUDA(cudaMemcpyAsync(best, d + o_best, np * 4, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    return ORBX_OK;
}

extern "C" int trx_triangulate_batch(trx_handle h, const float* pts1, const float* pts2, const int32_t* counts, const uint8_t* select, int nprob,
                                     int cap, const trx_cameras* cams, int nhyp, double* X, uint8_t* front, int32_t* nfront, int32_t* best)
{
    return triangulate_host<float>(h, pts1, pts2, counts, select, nprob, cap, cams, nhyp, X, front, nfront, best, "trx_triangulate_batch");
}

extern "C" int trx_triangulate(trx_handle h, const double* pts1, const double* pts2, int n, const double* Rt1, const double* Rt2, const double* K1,
                               const double* K2, double* X, uint8_t* front, int32_t* nfront)
{
    ORBX_REQUIRE(h != nullptr && n >= 0 && nfront != nullptr, "trx_triangulate: bad arguments");
    *nfront = 0;
    if (n == 0) return ORBX_OK;
    ORBX_REQUIRE(pts1 && pts2 && Rt1 && Rt2 && K1 && K2 && X, "trx_triangulate: NULL pointer");
    trx_cameras cam;
    memcpy(cam.Rt1, Rt1, sizeof(cam.Rt1)); memcpy(cam.Rt2, Rt2, sizeof(cam.Rt2));
    memcpy(cam.K1, K1, sizeof(cam.K1)); memcpy(cam.K2, K2, sizeof(cam.K2));
    const int32_t cnt = n;
    return triangulate_host<double>(h, pts1, pts2, &cnt, nullptr, 1, n, &cam, 1, X, front, nfront, nullptr, "trx_triangulate");
}

extern "C" int trx_triangulate_hypotheses(trx_handle h, const double* pts1, const double* pts2, int n, const double* Rt1, const double* Rts, int nhyp,
                                          const double* K1, const double* K2, double* X, int32_t* counts, int32_t* best)
{
    ORBX_REQUIRE(h != nullptr && n >= 0 && nhyp >= 1 && nhyp <= 64 && counts && best, "trx_triangulate_hypotheses: bad arguments");
    *best = -1;
    for (int i = 0; i < nhyp; i++) counts[i] = 0;
    ORBX_REQUIRE(Rt1 && Rts && K1 && K2, "trx_triangulate_hypotheses: NULL pointer");
    if (n == 0) { *best = 0; return ORBX_OK; }     // every count is 0: `maxCount < count` takes the first hypothesis (:344)
    ORBX_REQUIRE(pts1 && pts2, "trx_triangulate_hypotheses: NULL pointer");
    std::vector<trx_cameras> cams((size_t)nhyp);
    for (int i = 0; i < nhyp; i++) {
        memcpy(cams[i].Rt1, Rt1, 96); memcpy(cams[i].Rt2, Rts + 12 * i, 96);
        memcpy(cams[i].K1, K1, 72); memcpy(cams[i].K2, K2, 72);
    }
    std::vector<double> Xall((size_t)nhyp * n * 3);
    const int32_t cnt = n;
    int rc = triangulate_host<double>(h, pts1, pts2, &cnt, nullptr, 1, n, cams.data(), nhyp, Xall.data(), nullptr, counts, best,
                                      "trx_triangulate_hypotheses");
    if (rc) return rc;
    if (X && *best >= 0) memcpy(X, Xall.data() + (size_t)*best * n * 3, (size_t)n * 24);
    return ORBX_OK;
}

extern "C" int trx_associate_dev(trx_handle h, const orbx_dmatch* d_good, const int64_t* d_ngood, const uint8_t* d_status, const int32_t* d_premap,
                                 const int32_t* d_ncur, int nprob, int back, int cap, int32_t* d_cur_map, int32_t* d_assoc_q, int32_t* d_assoc_mp,
                                 int32_t* d_nassoc)
{
    ORBX_REQUIRE(h != nullptr, "trx_associate_dev: NULL handle");
    ORBX_REQUIRE(nprob >= 0 && back >= 1 && back <= 64 && cap >= 1 && (long long)back * cap < (1ll << 31), "trx_associate_dev: bad sizes");
    if (nprob == 0) return ORBX_OK;
    ORBX_REQUIRE(d_good && d_ngood && d_premap && d_ncur && d_cur_map && d_assoc_q && d_assoc_mp && d_nassoc, "trx_associate_dev: NULL pointer");
    ORBX_CUDA(cudaSetDevice(h->device));
    k_associate<<<nprob, AS_THREADS, 0, h->stream>>>(d_good, reinterpret_cast<const long long*>(d_ngood), d_status, d_premap, d_ncur, back, cap,
                                                     d_cur_map, d_assoc_q, d_assoc_mp, d_nassoc);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int trx_select_new_dev(trx_handle h, const orbx_dmatch* d_good, const int64_t* d_ngood, const uint8_t* d_status, int32_t* d_premap,
                                  int32_t* d_cur_map, const int32_t* d_ncur, const int32_t* d_next_id, int nprob, int back, int cap,
                                  uint8_t* d_accept, int32_t* d_nnew)
{
    ORBX_REQUIRE(h != nullptr, "trx_select_new_dev: NULL handle");
    ORBX_REQUIRE(nprob >= 0 && back >= 1 && back <= 64 && cap >= 1 && (long long)back * cap < (1ll << 31), "trx_select_new_dev: bad sizes");
    if (nprob == 0) return ORBX_OK;
    ORBX_REQUIRE(d_good && d_ngood && d_premap && d_cur_map && d_ncur && d_accept && d_nnew, "trx_select_new_dev: NULL pointer");
    ORBX_CUDA(cudaSetDevice(h->device));
    const size_t words = (size_t)nprob * cap * ((size_t)back + 1);
    int rc = trx_grow((void**)&h->d_scratch, &h->scratch_bytes, words * sizeof(uint32_t));
    if (rc) return rc;
    k_select_new<<<nprob, AS_THREADS, 0, h->stream>>>(d_good, reinterpret_cast<const long long*>(d_ngood), d_status, d_premap, d_cur_map, d_ncur,
                                                      d_next_id, back, cap, d_accept, h->d_scratch, h->d_scratch + (size_t)nprob * cap, d_nnew);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}
