// fmat.cu -- fundamental-matrix outlier filter for batches of matched frame pairs (sm_100a).
//
// Replaces computeFundamentalMatrix (reference src/CameraPoseEstimator.cpp:545-586), the step that follows matchFeatures
// for every matched frame pair (:291, :419):
//     F = findFundamentalMat(inputs1, inputs2, CV_FM_RANSAC, MAX_DISTANCE, CONFIDENCE, status)     (:563)
//     F = findFundamentalMat(inliers1, inliers2, CV_FM_8POINT)                                     (:585)
// OpenCV's RANSAC is sequential: sample 7 points, solve the cubic (<= 3 candidate matrices), count inliers, keep the best,
// shrink the iteration budget.  The result only depends on (a) the sample sequence, which is a function of the random
// generator and the points alone, and (b) the inlier count of every candidate.  So one CTA per pair works in rounds of up to
// 128 iterations:
//   1. one thread draws the index tuples of the round (cv::RNG, duplicate-free draws); the collinearity rejection is tested
//      for all samples in parallel, and the rare rejected sample makes one thread redo the rest of the round sequentially;
//   2. the samples are solved in parallel (one thread per sample; Householder null space + cv::solveCubic, double precision);
//   3. the candidates are scored one per warp, handed out in sequence order (single-precision classification with rigorous
//      error bounds, double precision for the few points near the threshold -- the decision is always OpenCV's);
//   4. warp 0 evaluates OpenCV's sequential best-update / iteration-budget rule over the round with two prefix scans,
// until the budget is exhausted -- same status mask as the sequential loop.  Candidates that cannot beat the best count
// completed before them are abandoned early, candidates beyond the budget that count implies are skipped (neither can be
// reached / chosen by the sequential loop).  The CTA then marks the inliers of the winning matrix and runs the normalised
// 8-point algorithm on them (block reduction of the 9x9 normal matrix, round-robin warp Jacobi for its smallest eigenvector,
// rank-2 projection).
//
// Double-precision arithmetic is IEEE, compiled without FMA contraction (build.sh: -fmad=false) like the x86 baseline build of
// OpenCV, so candidate matrices agree with the CPU to rounding of the transcendental calls and the inlier tests are the same.
#include "common.cuh"

#include <float.h>
#include <limits.h>
#include <stdlib.h>
#include <algorithm>

namespace orbx {
namespace {

// CTA size: 256 threads by default (2 pairs per SM), 128 as a tuning alternative (4 pairs per SM); see fm_launch.
constexpr int FM_MAXWARPS = 8;
constexpr int FM_MAXCHUNK = 128;      // iterations solved + scored per round
#ifndef FM_FIRSTCHUNK_N
#define FM_FIRSTCHUNK_N 64
#endif
#ifndef FM_CHUNKDIV
#define FM_CHUNKDIV 2
#endif
constexpr int FM_FIRSTCHUNK = FM_FIRSTCHUNK_N;     // iterations of the first round (64 / half the remaining budget measured best: profiles/README.md)
constexpr int FM_MODEL_POINTS = 7;
#ifndef FM_PPL_N
#define FM_PPL_N 4
#endif
constexpr int FM_PPL = FM_PPL_N;            // points per lane and scoring step (independent dependency chains)
constexpr int FM_SMEM_POINTS = 9000;  // pairs with at most this many correspondences keep them in shared memory (16 B each)

#ifdef FM_PROFILE
__device__ unsigned long long g_fm_prof[8];     // cycles of thread 0 per phase, summed over CTAs (debug builds only)
#define FM_TICK(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_fm_prof[k], (unsigned long long)(now_ - tick_)); tick_ = now_; } } while (0)
constexpr int FM_PROF_PAIRS = 4096;
__device__ unsigned long long g_fm_pair_ns[FM_PROF_PAIRS][2];    // %globaltimer when a pair's CTA starts and when it ends (every exit)
__device__ __forceinline__ unsigned long long fm_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
struct FmPairTimer {
    int pair;
    __device__ explicit FmPairTimer(int p) : pair(p) { if (threadIdx.x == 0 && pair < FM_PROF_PAIRS) g_fm_pair_ns[pair][0] = fm_globaltimer(); }
    __device__ ~FmPairTimer() { if (threadIdx.x == 0 && pair < FM_PROF_PAIRS) g_fm_pair_ns[pair][1] = fm_globaltimer(); }
};
#else
#define FM_TICK(k) do { } while (0)
#endif

constexpr int FM_EIGHT_WS = 88;          // doubles per pair between k_fm_ransac and k_fm_eight_point: A[81], c1x c1y s1 c2x c2y s2, valid

struct FmShared {
    double models[FM_MAXCHUNK][27];
    double best[9];
    double red[FM_MAXWARPS][48];
    double A[81], V[81];
    double rot_cs[4], rot_sn[4];      // the 4 concurrent Jacobi rotations of one round
    unsigned long long rng;
    unsigned long long rng_at[FM_MAXCHUNK];   // generator state before each sample of the chunk (restart point for the slow path)
    int idx[FM_MAXCHUNK][FM_MODEL_POINTS];
    int nmodels[FM_MAXCHUNK];
    int count[FM_MAXCHUNK][3];
    int rot_p[4], rot_q[4];
    int cmax[4];                      // largest |x1| |y1| |x2| |y2| of the pair (float bits; non-negative floats order like ints)
    int iter, niters, maxgood, chunk, chunk_redo, stop, have, next, first_bad, run_max, iter_limit;
    int min_median;                   // LMedS: float bits of the smallest median so far (+inf bits: none)
    float lmeds_t2;                   // LMedS: squared inlier threshold derived from it
};

__device__ __forceinline__ unsigned rng_next(unsigned long long& s)
{
    s = (unsigned long long)(unsigned)s * 4164903690ULL + (unsigned)(s >> 32);     // cv::RNG, multiply with carry
    return (unsigned)s;
}

__device__ bool have_collinear(const float2* p)
{
    const int i = FM_MODEL_POINTS - 1;
    for (int j = 0; j < i; j++) {
        const double dx1 = (double)p[j].x - (double)p[i].x, dy1 = (double)p[j].y - (double)p[i].y;
        for (int k = 0; k < j; k++) {
            const double dx2 = (double)p[k].x - (double)p[i].x, dy2 = (double)p[k].y - (double)p[i].y;
            if (fabs(dx2 * dy1 - dy2 * dx1) <= (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
        }
    }
    return false;
}

// x % n with a precomputed m = floor(2^32 / n): the quotient estimate is at most one too small
__device__ __forceinline__ int fast_mod(unsigned x, unsigned n, unsigned m)
{
    unsigned r = x - __umulhi(x, m) * n;
    if (r >= n) r -= n;
    if (r >= n) r -= n;
    return (int)r;
}

// the index draws of getSubset alone (7 distinct indices); the collinearity test is done in parallel afterwards
__device__ __forceinline__ void draw_indices(int n, unsigned m, unsigned long long& rng, int* out)
{
    int idx[FM_MODEL_POINTS];
#pragma unroll
    for (int i = 0; i < FM_MODEL_POINTS; i++) {
        for (;;) {
            const int v = n == 1 ? 0 : fast_mod(rng_next(rng), (unsigned)n, m);
            bool dup = false;
#pragma unroll
            for (int j = 0; j < i; j++) dup |= v == idx[j];
            if (!dup) { idx[i] = v; break; }
        }
    }
#pragma unroll
    for (int i = 0; i < FM_MODEL_POINTS; i++) out[i] = idx[i];
}

// RANSACPointSetRegistrator::getSubset: 7 distinct indices whose last point is not collinear with two earlier ones
__device__ bool get_subset(const float2* P1, const float2* P2, int n, unsigned long long& rng, int* out, int max_attempts)
{
    int idx[FM_MODEL_POINTS];
    float2 a[FM_MODEL_POINTS], b[FM_MODEL_POINTS];
    for (int iters = 0; iters < max_attempts; iters++) {
        for (int i = 0; i < FM_MODEL_POINTS;) {
            const int v = (int)(rng_next(rng) % (unsigned)n);
            int j;
            for (j = 0; j < i; j++)
                if (v == idx[j]) break;
            if (j < i) continue;
            idx[i] = v; a[i] = P1[v]; b[i] = P2[v];
            i++;
        }
        if (have_collinear(a) || have_collinear(b)) continue;
        for (int i = 0; i < FM_MODEL_POINTS; i++) out[i] = idx[i];
        return true;
    }
    return false;
}

// cv::solveCubic, real roots in OpenCV's order
__device__ int solve_cubic(const double* c, double* x)
{
    double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
    if (a0 == 0) {
        if (a1 == 0) {
            if (a2 == 0) return a3 == 0 ? -1 : 0;
            x[0] = -a3 / a2;
            return 1;
        }
        double d = a2 * a2 - 4 * a1 * a3;
        if (d < 0) return 0;
        d = sqrt(d);
        const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (fabs(q1) > fabs(q2)) { x[0] = q1 / a1; x[1] = a3 / q1; }
        else { x[0] = q2 / a1; x[1] = a3 / q2; }
        return d > 0 ? 2 : 1;
    }
    a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    const double R = (a1 * (2 * a1 * a1 - 9 * a2) + 27 * a3) * (1. / 54);
    const double Qcubed = Q * Q * Q;
    const double d = Qcubed - R * R;
    if (d > 0) {
        const double theta = acos(R / sqrt(Qcubed)), sqrtQ = sqrt(Q);
        const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
        x[0] = t0 * cos(t1) - t2;
        x[1] = t0 * cos(t1 + (2. * 3.14159265358979323846 / 3)) - t2;
        x[2] = t0 * cos(t1 + (4. * 3.14159265358979323846 / 3)) - t2;
        return 3;
    }
    if (d == 0) {
        if (R >= 0) { x[0] = -2 * pow(R, 1. / 3) - a1 / 3; x[1] = pow(R, 1. / 3) - a1 / 3; }
        else { x[0] = 2 * pow(-R, 1. / 3) - a1 / 3; x[1] = -pow(-R, 1. / 3) - a1 / 3; }
        return 2;
    }
    double e = pow(sqrt(-d) + fabs(R), 1. / 3);
    if (R > 0) e = -e;
    x[0] = (e + Q / e) - a1 * (1. / 3);
    return 1;
}

// F <- T2' F T1 (T = [s 0 -s cx; 0 s -s cy; 0 0 1]), then scaled to F[8] = 1 when |F[8]| > FLT_EPSILON
__device__ void denormalise(double* F, double c1x, double c1y, double s1, double c2x, double c2y, double s2)
{
    const double T1[9] = {s1, 0, -s1 * c1x, 0, s1, -s1 * c1y, 0, 0, 1}, T2[9] = {s2, 0, -s2 * c2x, 0, s2, -s2 * c2y, 0, 0, 1};
    double G[9], H[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) s += T2[k * 3 + i] * F[k * 3 + j];
            G[i * 3 + j] = s;
        }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) s += G[i * 3 + k] * T1[k * 3 + j];
            H[i * 3 + j] = s;
        }
    double t = 1.;
    if (fabs(H[8]) > (double)FLT_EPSILON) t = 1. / H[8];
#pragma unroll
    for (int i = 0; i < 9; i++) F[i] = H[i] * t;
}

// run7Point: candidate matrices of one 7-point sample
__device__ int solve7(const float2* P1, const float2* P2, const int* idx, double* models)
{
    double m[9][7], v[7][9];       // transposed system and the Householder vectors (local memory)
    double c1x = 0, c1y = 0, c2x = 0, c2y = 0, s1 = 0, s2 = 0;
    float2 a[7], b[7];
    for (int i = 0; i < 7; i++) {
        a[i] = P1[idx[i]]; b[i] = P2[idx[i]];
        c1x += a[i].x; c1y += a[i].y; c2x += b[i].x; c2y += b[i].y;
    }
    const double t = 1. / 7;
    c1x *= t; c1y *= t; c2x *= t; c2y *= t;
    for (int i = 0; i < 7; i++) {
        const double dx1 = a[i].x - c1x, dy1 = a[i].y - c1y, dx2 = b[i].x - c2x, dy2 = b[i].y - c2y;
        s1 += sqrt(dx1 * dx1 + dy1 * dy1);
        s2 += sqrt(dx2 * dx2 + dy2 * dy2);
    }
    s1 *= t; s2 *= t;
    if (s1 < (double)FLT_EPSILON || s2 < (double)FLT_EPSILON) return 0;
    s1 = sqrt(2.) / s1; s2 = sqrt(2.) / s2;
    for (int i = 0; i < 7; i++) {
        const double x0 = (a[i].x - c1x) * s1, y0 = (a[i].y - c1y) * s1, x1 = (b[i].x - c2x) * s2, y1 = (b[i].y - c2y) * s2;
        m[0][i] = x1 * x0; m[1][i] = x1 * y0; m[2][i] = x1;
        m[3][i] = y1 * x0; m[4][i] = y1 * y0; m[5][i] = y1;
        m[6][i] = x0; m[7][i] = y0; m[8][i] = 1;
    }
    // Householder QR of the 9x7 transposed system; the last two columns of Q span the null space
    for (int k = 0; k < 7; k++) {
        double norm = 0;
        for (int i = k; i < 9; i++) norm += m[i][k] * m[i][k];
        norm = sqrt(norm);
        for (int i = 0; i < 9; i++) v[k][i] = 0;
        if (norm == 0) continue;
        const double alpha = m[k][k] > 0 ? -norm : norm;
        for (int i = k; i < 9; i++) v[k][i] = m[i][k];
        v[k][k] -= alpha;
        double vn = 0;
        for (int i = k; i < 9; i++) vn += v[k][i] * v[k][i];
        if (vn == 0) continue;
        vn = 1. / sqrt(vn);
        for (int i = k; i < 9; i++) v[k][i] *= vn;
        for (int c = k; c < 7; c++) {
            double dot = 0;
            for (int i = k; i < 9; i++) dot += v[k][i] * m[i][c];
            for (int i = k; i < 9; i++) m[i][c] -= 2 * dot * v[k][i];
        }
    }
    double f1[9], f2[9];
    for (int j = 0; j < 2; j++) {
        double* q = j ? f2 : f1;
        for (int i = 0; i < 9; i++) q[i] = i == 7 + j;
        for (int k = 6; k >= 0; k--) {
            double dot = 0;
            for (int i = k; i < 9; i++) dot += v[k][i] * q[i];
            for (int i = k; i < 9; i++) q[i] -= 2 * dot * v[k][i];
        }
    }
    // F = lambda f1 + (1 - lambda) f2, det F = 0: cubic in lambda
    for (int i = 0; i < 9; i++) f1[i] -= f2[i];
    double c[4], r[3] = {0, 0, 0};
    double t0 = f2[4] * f2[8] - f2[5] * f2[7], t1 = f2[3] * f2[8] - f2[5] * f2[6], t2 = f2[3] * f2[7] - f2[4] * f2[6];
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
           f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
           f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
           f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    t0 = f1[4] * f1[8] - f1[5] * f1[7]; t1 = f1[3] * f1[8] - f1[5] * f1[6]; t2 = f1[3] * f1[7] - f1[4] * f1[6];
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
           f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
           f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
           f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    const int n = solve_cubic(c, r);
    if (n < 1 || n > 3) return 0;
    for (int k = 0; k < n; k++) {
        double F[9];
        double lambda = r[k], mu = 1.;
        const double s = f1[8] * r[k] + f2[8];
        if (fabs(s) > DBL_EPSILON) { mu = 1. / s; lambda *= mu; F[8] = 1.; }
        else F[8] = 0.;
        for (int i = 0; i < 8; i++) F[i] = f1[i] * lambda + f2[i] * mu;
        denormalise(F, c1x, c1y, s1, c2x, c2y, s2);
        for (int i = 0; i < 9; i++) models[9 * k + i] = F[i];
    }
    return n;
}

// FMEstimatorCallback::computeError for one correspondence, then findInliers' test `(float)max(d1*d1*s1, d2*d2*s2) <= t2`
// with s = 1/(a*a + b*b), exactly as OpenCV evaluates it (Fm: the matrix in shared memory).
__device__ __noinline__ float epi_error(const double* Fm, float2 p1, float2 p2)
{
    const double x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
    double a = Fm[0] * x1 + Fm[1] * y1 + Fm[2], b = Fm[3] * x1 + Fm[4] * y1 + Fm[5], c = Fm[6] * x1 + Fm[7] * y1 + Fm[8];
    const double s2 = 1. / (a * a + b * b), d2 = x2 * a + y2 * b + c;
    a = Fm[0] * x2 + Fm[3] * y2 + Fm[6]; b = Fm[1] * x2 + Fm[4] * y2 + Fm[7]; c = Fm[2] * x2 + Fm[5] * y2 + Fm[8];
    const double s1 = 1. / (a * a + b * b), d1 = x1 * a + y1 * b + c;
    const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
    return __double2float_rn(e1 < e2 ? e2 : e1);             // (float)std::max(e1, e2)
}

__device__ __forceinline__ bool is_inlier_exact(const double* Fm, float2 p1, float2 p2, float t2) { return epi_error(Fm, p1, p2) <= t2; }

// The two divisions are only needed for points within 1e-6 (relative) of the threshold: with the same a, b, d as OpenCV
// computes, q = d*d <= t2*(1 - 1e-9)*den implies q*fl(1/den) < t2, and q >= t2*(1 + 1e-6)*den implies it exceeds the largest
// double that still rounds to the float t2.  Branch-free: 1 inlier, 0 outlier, -1 undecided (-> is_inlier_exact).
__device__ __forceinline__ int classify(const double* F, float2 p1, float2 p2, double tlo, double thi)
{
    const double x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
    double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
    const double den2 = a * a + b * b, d2 = x2 * a + y2 * b + c;
    a = F[0] * x2 + F[3] * y2 + F[6]; b = F[1] * x2 + F[4] * y2 + F[7]; c = F[2] * x2 + F[5] * y2 + F[8];
    const double den1 = a * a + b * b, d1 = x1 * a + y1 * b + c;
    const double q1 = d1 * d1, q2 = d2 * d2;
    const bool regular = den1 > 0 && den1 < 1e300 && den2 > 0 && den2 < 1e300;
    const bool in = q1 <= tlo * den1 && q2 <= tlo * den2;
    const bool out = q1 >= thi * den1 || q2 >= thi * den2;
    return regular && (in || out) ? (in ? 1 : 0) : -1;
}

// Single-precision pre-classification of the same test, with rigorous error bounds (u = 2^-24):
//   a_f = fma(F0f, x, fma(F1f, y, F2f)) differs from the value OpenCV computes in double by at most 3u(|F0 x| + |F1 y| + |F2|);
//   d_f = fma(x', a_f, fma(y', b_f, c_f)) by at most 5u(|x'| A + |y'| B + C), A, B, C being those magnitude sums.
// The magnitude sums are bounded once per candidate with the largest coordinates of the pair (Cand32), with 4u / 8u instead of
// 3u / 5u.  A point is decided here only if it is inside (outside) for every value within the bounds, against thresholds moved
// by a further 4e-6 (covers the roundings of the bound arithmetic itself); otherwise -1 -> the double-precision classify.
struct Cand32 {
    float F[9];
    float ea2, eb2, E2;      // side 2: line F p1, distance of p2
    float ea1, eb1, E1;      // side 1: line F' p2, distance of p1
};

__device__ __forceinline__ void cand32_init(Cand32& c, const double* F, const float* cmax /* X1 Y1 X2 Y2 */)
{
#pragma unroll
    for (int i = 0; i < 9; i++) c.F[i] = __double2float_rn(F[i]);
    const float X1 = cmax[0], Y1 = cmax[1], X2 = cmax[2], Y2 = cmax[3], tiny = 1e-30f;
    const float u4 = 2.384185791015625e-07f, u8 = 4.76837158203125e-07f;       // 4 * 2^-24, 8 * 2^-24
    const float A2 = fmaf(fabsf(c.F[0]), X1, fmaf(fabsf(c.F[1]), Y1, fabsf(c.F[2]))) + tiny;
    const float B2 = fmaf(fabsf(c.F[3]), X1, fmaf(fabsf(c.F[4]), Y1, fabsf(c.F[5]))) + tiny;
    const float C2 = fmaf(fabsf(c.F[6]), X1, fmaf(fabsf(c.F[7]), Y1, fabsf(c.F[8]))) + tiny;
    c.ea2 = u4 * A2; c.eb2 = u4 * B2; c.E2 = u8 * fmaf(X2, A2, fmaf(Y2, B2, C2));
    const float A1 = fmaf(fabsf(c.F[0]), X2, fmaf(fabsf(c.F[3]), Y2, fabsf(c.F[6]))) + tiny;
    const float B1 = fmaf(fabsf(c.F[1]), X2, fmaf(fabsf(c.F[4]), Y2, fabsf(c.F[7]))) + tiny;
    const float C1 = fmaf(fabsf(c.F[2]), X2, fmaf(fabsf(c.F[5]), Y2, fabsf(c.F[8]))) + tiny;
    c.ea1 = u4 * A1; c.eb1 = u4 * B1; c.E1 = u8 * fmaf(X1, A1, fmaf(Y1, B1, C1));
}

// one side: 1 surely inside, 0 surely outside, -1 undecided
__device__ __forceinline__ int side32(float a, float b, float d, float ea, float eb, float E, float tlo32, float thi32)
{
    const float la = fmaxf(fabsf(a) - ea, 0.f), lb = fmaxf(fabsf(b) - eb, 0.f), ha = fabsf(a) + ea, hb = fabsf(b) + eb;
    const float den_lo = fmaf(la, la, lb * lb), den_hi = fmaf(ha, ha, hb * hb);
    const float dhi = fabsf(d) + E, dlo = fmaxf(fabsf(d) - E, 0.f);
    const bool regular = den_lo > 1e-30f && den_hi < 1e30f && dhi < 1e15f;        // no overflow / underflow anywhere above
    const bool in = dhi * dhi <= tlo32 * den_lo, out = dlo * dlo >= thi32 * den_hi;
    return regular ? (in ? 1 : (out ? 0 : -1)) : -1;
}

__device__ __forceinline__ int classify32(const Cand32& c, float2 p1, float2 p2, float tlo32, float thi32)
{
    const float x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
    const float a2 = fmaf(c.F[0], x1, fmaf(c.F[1], y1, c.F[2])), b2 = fmaf(c.F[3], x1, fmaf(c.F[4], y1, c.F[5]));
    const float c2 = fmaf(c.F[6], x1, fmaf(c.F[7], y1, c.F[8])), d2 = fmaf(x2, a2, fmaf(y2, b2, c2));
    const float a1 = fmaf(c.F[0], x2, fmaf(c.F[3], y2, c.F[6])), b1 = fmaf(c.F[1], x2, fmaf(c.F[4], y2, c.F[7]));
    const float c1 = fmaf(c.F[2], x2, fmaf(c.F[5], y2, c.F[8])), d1 = fmaf(x1, a1, fmaf(y1, b1, c1));
    const int s2 = side32(a2, b2, d2, c.ea2, c.eb2, c.E2, tlo32, thi32), s1 = side32(a1, b1, d1, c.ea1, c.eb1, c.E1, tlo32, thi32);
    // inlier needs both sides inside; one side outside is enough for an outlier
    return (s1 == 0 || s2 == 0) ? 0 : ((s1 == 1 && s2 == 1) ? 1 : -1);
}

// double-precision decision for one point: division-free classification, then OpenCV's own expression (F: the matrix in
// registers, or NULL to read it from Fm)
__device__ __forceinline__ bool is_inlier(const double* F, const double* Fm, float2 p1, float2 p2, float t2, double tlo, double thi)
{
    const int c = classify(F ? F : Fm, p1, p2, tlo, thi);
    return c >= 0 ? c != 0 : is_inlier_exact(Fm, p1, p2, t2);
}

// min(n0, RANSACUpdateNumIters(p, (n - good) / n, .)): the iteration budget after a candidate with `good` inliers became the best,
// given a budget n0 before.  lognum = log(max(1 - p, DBL_MIN)).  The budget formula log(1 - p) / log(1 - w^7), w = good / n, is far
// above n0 for most candidates; a single-precision estimate (relative error < 1e-2 wherever it is used to decide) screens those
// out, only estimates below 2 n0 + 2 take OpenCV's double-precision expression.
__device__ int iter_limit_of(double lognum, int n, int good, int n0)
{
    const float w = (float)good / (float)n, w2 = w * w, w7 = w2 * w2 * w2 * w, df = 1.f - w7;
    const float lim = 2.f * (float)n0 + 2.f;
    if (df >= 1.f) {
        // w^7 <= 3e-8 is below single-precision resolution next to 1: the formula is at least |lognum| / 3.1e-8, which only
        // decides the matter for confidences that are not tiny; otherwise fall through to the double-precision expression
        if ((float)(-lognum) >= lim * 1.2e-7f) return n0;
    } else {
        const float qf = (float)lognum / logf(df);                  // df == 0: -0 -> not screened out
        if (!(qf < lim)) return n0;                                 // df is within a factor 1.5 of 1 - w^7 in its log: 2 n0 suffices
    }
    double ep = (double)(n - good) / n;
    ep = fmin(fmax(ep, 0.), 1.);
    double denom = 1. - pow(1. - ep, (double)FM_MODEL_POINTS);
    if (denom < DBL_MIN) return 0;
    denom = log(denom);
    if (denom >= 0) return n0;
    const double q = lognum / denom;
    return q >= (double)n0 ? n0 : min(n0, __double2int_rn(q));
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sums `NV` per-thread values over the CTA; the totals land in sh.red[0][0..NV)
template <int NV, int FM_WARPS>
__device__ void block_sum(FmShared& sh, const double* v)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const double s = warp_sum(v[i]);
        if (lane == 0) sh.red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < FM_WARPS; w++) s += sh.red[w][threadIdx.x];
        sh.red[0][threadIdx.x] = s;
    }
    __syncthreads();
}

// Jacobi on the symmetric 9x9 sh.A by warp 0; eigenvectors = rows of sh.V.  Round-robin ordering: each of the 9 rounds of a
// sweep applies 4 rotations on disjoint index pairs at once -- lanes 0..3 compute the angles, then 36 lanes update the
// 4 x 9 column pairs, then the row pairs (and V).
template <class Shared>
__device__ void jacobi9_warp(Shared& sh, int lane)
{
    for (int i = lane; i < 81; i += 32) sh.V[i] = (i / 9 == i % 9) ? 1. : 0.;
    __syncwarp();
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0, diag = 0;
        for (int i = lane; i < 81; i += 32) {
            const double x = sh.A[i] * sh.A[i];
            if (i / 9 == i % 9) diag += x; else off += x;
        }
        off = warp_sum(off); diag = warp_sum(diag);
        if (off <= 1e-31 * diag) break;                    // off-diagonal entries at the rounding floor of the diagonal
        for (int r = 0; r < 9; r++) {
            if (lane < 4) {
                const int a = (r + lane + 1) % 9, b = (r + 8 - lane) % 9;
                const int p = min(a, b), q = max(a, b);
                const double apq = sh.A[p * 9 + q];
                double cs = 1., sn = 0.;
                if (apq != 0) {
                    // The tangent only has to be good enough for the iteration to converge (an error of 1e-7 leaves that much of
                    // a_pq for the next sweep), so it is computed in single precision; the rotation itself is orthogonal to
                    // double precision because cs and sn are derived from that one t in double.
                    const float theta = (float)(sh.A[q * 9 + q] - sh.A[p * 9 + p]) / (2.f * (float)apq);
                    const float tf = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(fmaf(theta, theta, 1.f)));
                    const double t = (theta == theta) ? (double)tf : 0.;           // 0/0 or inf/inf in float: skip this pair
                    cs = rsqrt(t * t + 1);
                    sn = t * cs;
                }
                sh.rot_p[lane] = p; sh.rot_q[lane] = q; sh.rot_cs[lane] = cs; sh.rot_sn[lane] = sn;
            }
            __syncwarp();
#pragma unroll
            for (int t = lane; t < 36; t += 32) {
                const int k = t / 9, i = t - 9 * k, p = sh.rot_p[k], q = sh.rot_q[k];
                const double cs = sh.rot_cs[k], sn = sh.rot_sn[k];
                const double akp = sh.A[i * 9 + p], akq = sh.A[i * 9 + q];
                sh.A[i * 9 + p] = cs * akp - sn * akq;
                sh.A[i * 9 + q] = sn * akp + cs * akq;
            }
            __syncwarp();
#pragma unroll
            for (int t = lane; t < 36; t += 32) {
                const int k = t / 9, i = t - 9 * k, p = sh.rot_p[k], q = sh.rot_q[k];
                const double cs = sh.rot_cs[k], sn = sh.rot_sn[k];
                const double apk = sh.A[p * 9 + i], aqk = sh.A[q * 9 + i];
                sh.A[p * 9 + i] = cs * apk - sn * aqk;
                sh.A[q * 9 + i] = sn * apk + cs * aqk;
                const double vpk = sh.V[p * 9 + i], vqk = sh.V[q * 9 + i];
                sh.V[p * 9 + i] = cs * vpk - sn * vqk;
                sh.V[q * 9 + i] = sn * vpk + cs * vqk;
            }
            __syncwarp();
        }
    }
    __syncwarp();
}

// Jacobi on a symmetric 3x3 (one thread); returns the eigenvector of the smallest eigenvalue in vs
__device__ void smallest_eigvec3(double* G, double* vs)
{
    double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int sweep = 0; sweep < 40; sweep++) {
        const double off = G[1] * G[1] + G[2] * G[2] + G[5] * G[5], diag = G[0] * G[0] + G[4] * G[4] + G[8] * G[8];
        if (off == 0 || off <= 1e-36 * diag) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                const double apq = G[p * 3 + q];
                if (apq == 0) continue;
                const double theta = (G[q * 3 + q] - G[p * 3 + p]) / (2 * apq);
                const double t = (theta >= 0 ? 1. : -1.) / (fabs(theta) + sqrt(theta * theta + 1));
                const double cs = 1. / sqrt(t * t + 1), sn = t * cs;
                for (int k = 0; k < 3; k++) {
                    const double akp = G[k * 3 + p], akq = G[k * 3 + q];
                    G[k * 3 + p] = cs * akp - sn * akq; G[k * 3 + q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; k++) {
                    const double apk = G[p * 3 + k], aqk = G[q * 3 + k];
                    G[p * 3 + k] = cs * apk - sn * aqk; G[q * 3 + k] = sn * apk + cs * aqk;
                    const double vpk = V[p * 3 + k], vqk = V[q * 3 + k];
                    V[p * 3 + k] = cs * vpk - sn * vqk; V[q * 3 + k] = sn * vpk + cs * vqk;
                }
            }
    }
    int m = 0;
    if (G[4] < G[m * 4]) m = 1;
    if (G[8] < G[m * 4]) m = 2;
    for (int k = 0; k < 3; k++) vs[k] = V[m * 3 + k];
}

// after jacobi9_warp: the eigenvector of the smallest eigenvalue as F, rank 2 enforced, normalisation undone (lane 0)
template <class Shared>
__device__ void eight_point_finish(Shared& sh, double c1x, double c1y, double s1, double c2x, double c2y, double s2, double* Fo, int32_t* inf)
{
    int small = 0, m = 0;
    for (int i = 0; i < 9; i++) {
        if (fabs(sh.A[i * 10]) < DBL_EPSILON) small++;
        if (sh.A[i * 10] < sh.A[m * 10]) m = i;
    }
    if (small < 2) {                                   // run8Point: the 8 largest eigenvalues must be non-zero
        double F0[9], G[9], vs[3];
        for (int i = 0; i < 9; i++) F0[i] = sh.V[m * 9 + i];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) {
                double s = 0;
                for (int c = 0; c < 3; c++) s += F0[c * 3 + a] * F0[c * 3 + b];
                G[a * 3 + b] = s;
            }
        smallest_eigvec3(G, vs);
        // rank 2: F0 - (F0 v) v' == U diag(w0, w1, 0) V'
        for (int a = 0; a < 3; a++) {
            const double fv = F0[a * 3] * vs[0] + F0[a * 3 + 1] * vs[1] + F0[a * 3 + 2] * vs[2];
            for (int b = 0; b < 3; b++) F0[a * 3 + b] -= fv * vs[b];
        }
        denormalise(F0, c1x, c1y, s1, c2x, c2y, s2);
        for (int i = 0; i < 9; i++) Fo[i] = F0[i];
        inf[3] = 1;
    }
}

// One CTA per pair.  pts1/pts2: npairs x cap correspondences (float x, y); counts[pair] of them valid.
// status: npairs x cap bytes (0/1); F: npairs x 9 doubles (zeros: no result); info: npairs x 4 ints
// {inliers, iterations run, candidates scored, 0}.
#ifndef FM_CTAS_PER_SM
#define FM_CTAS_PER_SM (512 / FM_THREADS)   // 128 registers per thread; 3 CTAs of 256 threads (80 registers) were measured: see profiles/README.md
#endif
template <bool SMEM_POINTS, int FM_THREADS>
__global__ void __launch_bounds__(FM_THREADS, FM_CTAS_PER_SM)
k_fm_ransac(const float2* __restrict__ pts1, const float2* __restrict__ pts2, const int32_t* __restrict__ counts, int cap,
            double thr, double conf, int max_iters, uint8_t* __restrict__ status, double* __restrict__ Fout, int32_t* __restrict__ info,
            double* __restrict__ eight_ws /* null: the 8-point step ends here; else [npairs][FM_EIGHT_WS] for k_fm_eight_point */)
{
    extern __shared__ __align__(16) unsigned char fm_smem[];
    FmShared& sh = *reinterpret_cast<FmShared*>(fm_smem);
    const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef FM_PROFILE
    FmPairTimer pair_timer_(pair);
#endif
    const int n = min(counts[pair], cap);
    const float2* P1 = pts1 + (size_t)pair * cap;
    const float2* P2 = pts2 + (size_t)pair * cap;
    uint8_t* st = status + (size_t)pair * cap;
    double* Fo = Fout + (size_t)pair * 9;
    int32_t* inf = info + (size_t)pair * 4;

    for (int i = tid; i < cap; i += FM_THREADS) st[i] = 0;
    if (tid < 9) Fo[tid] = 0.;
    if (tid < 4) inf[tid] = 0;
    if (eight_ws && tid == 0) eight_ws[(size_t)pair * FM_EIGHT_WS + 87] = 0.;        // no system for the second kernel (yet)
    // OpenCV: fewer than 7 points -> empty result; exactly 7 -> the bare 7-point solver (up to three stacked matrices the
    // reference cannot use): both report "no model" here.  8..14 points -> least median of squares instead of RANSAC
    // (findFundamentalMat: `npoints >= 15` selects RANSAC): same sampler, a fixed iteration count, the candidate with the
    // smallest median error wins.  (With fewer than 14 points that median is the error of a point the candidate was fitted
    // to, ~1e-25, so OpenCV's own choice is decided by rounding noise; the result is a valid LMedS answer, not necessarily
    // OpenCV's.  With 14 it is reproducible, and tested.)
    if (n < 8) return;
    const bool lmeds = n < 15;

    if (tid < 4) sh.cmax[tid] = 0;
    __syncthreads();
    {
        float2* s1 = reinterpret_cast<float2*>(fm_smem + sizeof(FmShared));
        float2* s2 = s1 + n;
        float mx[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = tid; i < n; i += FM_THREADS) {
            const float2 a = P1[i], b = P2[i];
            if (SMEM_POINTS) { s1[i] = a; s2[i] = b; }
            mx[0] = fmaxf(mx[0], fabsf(a.x)); mx[1] = fmaxf(mx[1], fabsf(a.y));
            mx[2] = fmaxf(mx[2], fabsf(b.x)); mx[3] = fmaxf(mx[3], fabsf(b.y));
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int o = 16; o; o >>= 1) mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
            if (lane == 0) atomicMax(&sh.cmax[k], __float_as_int(mx[k]));
        }
        if (SMEM_POINTS) { P1 = s1; P2 = s2; }
    }
    if (thr <= 0) thr = 3;
    if (conf < DBL_EPSILON || conf > 1 - DBL_EPSILON) conf = 0.99;
    float t2 = __double2float_rn(thr * thr);
    double tlo = (double)t2 * (1. - 1e-9), thi = (double)t2 * (1. + 1e-6);
    const float tlo32 = __double2float_rd(tlo * (1. - 4e-6)), thi32 = __double2float_ru(thi * (1. + 4e-6));
    const unsigned nmod = (unsigned)(0x100000000ULL / (unsigned)n);
    const double lognum = log(fmax(1. - fmin(fmax(conf, 0.), 1.), DBL_MIN));
    if (tid == 0) {
        sh.rng = 0xffffffffffffffffULL;
        sh.iter = 0; sh.niters = max_iters; sh.maxgood = 0; sh.stop = 0; sh.have = 0;
        sh.min_median = 0x7f800000;
        if (lmeds) {        // LMeDSPointSetRegistrator: RANSACUpdateNumIters(confidence, outlierRatio = 0.45, 7, maxIters), at least 3
            const double denom = log(1. - pow(1. - 0.45, (double)FM_MODEL_POINTS));
            const double q = lognum / denom;
            sh.niters = max(3, q >= (double)max_iters ? max_iters : __double2int_rn(q));
        }
    }
    __syncthreads();

#ifdef FM_PROFILE
    long long tick_ = clock64();
#endif
    int scored = 0;
    const float cm[4] = {__int_as_float(sh.cmax[0]), __int_as_float(sh.cmax[1]), __int_as_float(sh.cmax[2]), __int_as_float(sh.cmax[3])};
    for (int round = 0;; round++) {
        // ---- 1. samples of the next chunk of iterations: one thread draws the index tuples (integer work only), then every
        // sample's collinearity test runs in parallel.  A rejected sample (rare: three points exactly on a line, or
        // coincident points) changes what the generator produces next, so the chunk is redone from there by the
        // sequential rule.
        if (tid == 0) {
            // iterations per round: 1/FM_CHUNKDIV of the remaining budget, between FM_FIRSTCHUNK and FM_MAXCHUNK.  The budget
            // usually collapses as soon as a good sample is scored; candidates past that point are skipped by the running
            // iteration limit of step 3, so a long round costs little, and it has fewer serial sections than several short ones.
            const int remaining = sh.niters - sh.iter;
            const int chunk = lmeds ? min(remaining, FM_MAXCHUNK)
                                    : min(remaining, round == 0 ? FM_FIRSTCHUNK : min(FM_MAXCHUNK, max(FM_FIRSTCHUNK, remaining / FM_CHUNKDIV)));
            unsigned long long rng = sh.rng;
            for (int i = 0; i < chunk; i++) {
                sh.rng_at[i] = rng;
                draw_indices(n, nmod, rng, sh.idx[i]);
            }
            sh.rng = rng;
            sh.chunk = chunk;
            sh.next = 0;
            sh.first_bad = chunk;
            sh.run_max = max(sh.maxgood, FM_MODEL_POINTS - 1);
            sh.iter_limit = (sh.niters << 8) | 0xff;        // (budget << 8) | iteration of the candidate that implied it
        }
        FM_TICK(0);
        __syncthreads();
        if (tid < sh.chunk) {
            float2 a[FM_MODEL_POINTS], b[FM_MODEL_POINTS];
#pragma unroll
            for (int i = 0; i < FM_MODEL_POINTS; i++) { a[i] = P1[sh.idx[tid][i]]; b[i] = P2[sh.idx[tid][i]]; }
            if (have_collinear(a) || have_collinear(b)) atomicMin(&sh.first_bad, tid);
        }
        __syncthreads();
        // (thread 0 publishes the redone round's length in a separate word: the other threads may still be reading sh.chunk and
        // sh.first_bad for the branch below while it is already working)
        const int drawn = sh.chunk, first_bad = sh.first_bad;
        if (first_bad < drawn) {
            if (tid == 0) {
                int chunk = drawn;
                unsigned long long rng = sh.rng_at[first_bad];
                for (int i = first_bad; i < chunk; i++)
                    if (!get_subset(P1, P2, n, rng, sh.idx[i], lmeds ? 1000 : 10000)) { chunk = i; sh.stop = 1; break; }     // OpenCV leaves its loop here
                sh.rng = rng;
                sh.chunk_redo = chunk;
            }
            __syncthreads();
        }
        const int chunk = first_bad < drawn ? sh.chunk_redo : drawn;
        FM_TICK(1);
        if (chunk == 0) break;
        const int iter0 = sh.iter, niters0 = sh.niters;
        // ---- 2. one thread per sample: candidate matrices
        if (tid < chunk) sh.nmodels[tid] = solve7(P1, P2, sh.idx[tid], sh.models[tid]);
        __syncthreads();
        FM_TICK(2);
        if (lmeds) {
            // ---- 3'/4' (8..14 points): one thread per candidate computes OpenCV's float errors and their median (element n/2 of
            // the errors sorted as integers, as LMeDSPointSetRegistrator does); warp 0 then takes the first candidate in sequence
            // order whose median is below every earlier one.
            for (int m = tid; m < 3 * chunk; m += FM_THREADS) {
                const int it = m / 3, k = m - 3 * it;
                int med = 0x7fffffff;
                if (k < sh.nmodels[it]) {
                    const double* Fm = sh.models[it] + 9 * k;
                    int e[14];
                    for (int i = 0; i < n; i++) {
                        const int v = __float_as_int(epi_error(Fm, P1[i], P2[i]));
                        int j = i;
                        for (; j > 0 && e[j - 1] > v; j--) e[j] = e[j - 1];
                        e[j] = v;
                    }
                    med = e[n / 2];
                    scored++;
                }
                sh.count[it][k] = med;
            }
            __syncthreads();
            if (warp == 0) {
                unsigned long long best = ~0ULL;
#pragma unroll
                for (int e = 0; e < 12; e++) {
                    const int i = lane * 4 + e / 3;
                    if (i < chunk) {
                        const unsigned long long key = ((unsigned long long)(unsigned)sh.count[i][e % 3] << 32) | (unsigned)(lane * 12 + e);
                        best = min(best, key);
                    }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
                if (lane == 0) {
                    const int bits = (int)(best >> 32), g = (int)(best & 0xffffffffu);
                    if (bits >= 0 && bits < sh.min_median) {          // `median < minMedian`; NaN / inf medians never qualify
                        sh.min_median = bits;
                        const double* m = sh.models[g / 3] + 9 * (g % 3);
                        for (int j = 0; j < 9; j++) sh.best[j] = m[j];
                        sh.have = 1;
                    }
                    sh.iter = iter0 + chunk;
                    if (sh.iter >= sh.niters) sh.stop = 1;
                }
            }
            __syncthreads();
            if (sh.stop) break;
            continue;
        }
        // ---- 3. one warp per candidate, handed out in sequence order: inlier count.
        // Two facts a warp reads BEFORE it takes the next candidate let it do less without changing what the sequential rule
        // (step 4) will decide.  Every candidate completed by then precedes the one taken (candidates are handed out in order),
        // so (a) a count that cannot exceed sh.run_max -- the best exact count completed so far -- can never become the best:
        // the candidate is abandoned as soon as that is certain; (b) once a completed count c bounds the iteration budget by
        // iter_limit_of(c), a candidate of a later iteration at or beyond that budget is never reached and is skipped altogether.
        for (;;) {
            int m = 0, bound = 0, limit = 0;
            if (lane == 0) {
                bound = *(volatile int*)&sh.run_max;
                limit = *(volatile int*)&sh.iter_limit;
                __threadfence_block();
                m = atomicAdd(&sh.next, 1);
            }
            m = __shfl_sync(0xffffffffu, m, 0);
            bound = __shfl_sync(0xffffffffu, bound, 0);
            limit = __shfl_sync(0xffffffffu, limit, 0);
            if (m >= 3 * chunk) break;
            const int it = m / 3, k = m - 3 * it;
            if (k >= sh.nmodels[it]) continue;
            // the budget a completed candidate implies takes effect at the top of the NEXT iteration: it only rules out
            // candidates of strictly later iterations (the other models of its own iteration are still scored by OpenCV)
            if ((limit & 0xff) < it && iter0 + it >= (limit >> 8)) { if (lane == 0) sh.count[it][k] = 0; continue; }
            const double* Fm = sh.models[it] + 9 * k;
            Cand32 c32;
            cand32_init(c32, Fm, cm);
            int cnt = 0;
            for (int base = 0; base < n; base += 32 * FM_PPL) {       // FM_PPL independent points per lane and step
                int c[FM_PPL];
                float2 pa[FM_PPL], pb[FM_PPL];
#pragma unroll
                for (int r = 0; r < FM_PPL; r++) {
                    const int j = min(base + 32 * r + lane, n - 1);
                    pa[r] = P1[j]; pb[r] = P2[j];
                }
                int any = 0;
#pragma unroll
                for (int r = 0; r < FM_PPL; r++) { c[r] = classify32(c32, pa[r], pb[r], tlo32, thi32); any |= c[r]; }
                if (__any_sync(0xffffffffu, any < 0)) {                // rare: a point within ~1e-4 of the threshold
#pragma unroll
                    for (int r = 0; r < FM_PPL; r++)
                        if (c[r] < 0) c[r] = is_inlier(nullptr, Fm, pa[r], pb[r], t2, tlo, thi);
                }
#pragma unroll
                for (int r = 0; r < FM_PPL; r++) cnt += __popc(__ballot_sync(0xffffffffu, c[r] && base + 32 * r + lane < n));
                if (cnt + max(n - base - 32 * FM_PPL, 0) <= bound) { cnt = 0; break; }
            }
            if (lane == 0) {
                sh.count[it][k] = cnt;
                if (cnt > bound && cnt > atomicMax(&sh.run_max, cnt))
                    atomicMin(&sh.iter_limit, (iter_limit_of(lognum, n, cnt, niters0) << 8) | it);    // it < 128; smallest budget wins
            }
            scored++;
        }
        __syncthreads();
        FM_TICK(3);
        // ---- 4. OpenCV's sequential update rule over the chunk, evaluated by warp 0 in parallel.  In sequence order a
        // candidate becomes the best iff its count exceeds every earlier one (and the floor): an exclusive prefix maximum.
        // Each such improvement j lowers the iteration budget to min(budget, r_j), r_j = RANSACUpdateNumIters' closed form, and
        // the loop reaches it iff its iteration index is below the budget left by the improvements of EARLIER ITERATIONS (the
        // budget is checked once per iteration, not per model): an exclusive prefix minimum over iterations.  Budgets only shrink and indices only grow, so the improvements reached form a prefix; the last one
        // reached is the result of the round.
        if (warp == 0) {
            static_assert(FM_MAXCHUNK == 128, "4 iterations per lane");
            const int N0 = sh.niters, floor0 = max(sh.maxgood, FM_MODEL_POINTS - 1);
            int cnt[12], lim[12], lmax = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = lane * 4 + j;
                const int nm = i < chunk ? sh.nmodels[i] : 0;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int c = k < nm ? sh.count[i][k] : 0;
                    cnt[3 * j + k] = c;
                    lmax = max(lmax, c);
                }
            }
            int incl = lmax;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl = max(incl, v);
            }
            int run = __shfl_up_sync(0xffffffffu, incl, 1);
            run = max(lane ? run : 0, floor0);
            int lmin = INT_MAX;
#pragma unroll
            for (int e = 0; e < 12; e++) {
                lim[e] = INT_MAX;
                if (cnt[e] > run) {
                    run = cnt[e];
                    lim[e] = iter_limit_of(lognum, n, cnt[e], N0);
                    lmin = min(lmin, lim[e]);
                } else {
                    cnt[e] = -1;                                   // not an improvement
                }
            }
            int pmin = lmin;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, pmin, o);
                if (lane >= o) pmin = min(pmin, v);
            }
            int P = __shfl_up_sync(0xffffffffu, pmin, 1);
            P = min(lane ? P : INT_MAX, N0);
            // P: budget in force at the top of the iteration being looked at (improvements of strictly earlier iterations);
            // Q: budget including the improvements of this iteration seen so far -- OpenCV checks the budget once per iteration,
            // so the models of one iteration are all scored and may improve one after the other
            int last = -1, last_cnt = 0, last_n = 0, Q = P;
#pragma unroll
            for (int e = 0; e < 12; e++) {
                if (e % 3 == 0) P = Q;
                if (cnt[e] >= 0) {
                    if (iter0 + lane * 4 + e / 3 < P) { last = lane * 12 + e; last_cnt = cnt[e]; last_n = min(Q, lim[e]); }
                    Q = min(Q, lim[e]);
                }
            }
            int g = last;
#pragma unroll
            for (int o = 16; o; o >>= 1) g = max(g, __shfl_xor_sync(0xffffffffu, g, o));
            if (g >= 0 && g == last) {                             // the lane that owns the last improvement reached
                const double* m = sh.models[g / 3] + 9 * (g % 3);
                for (int j = 0; j < 9; j++) sh.best[j] = m[j];
                sh.have = 1;
                sh.maxgood = last_cnt;
                sh.niters = last_n;
            }
            __syncwarp();
            if (lane == 0) {
                const int nf = sh.niters;
                const int done = min(chunk, max(g >= 0 ? g / 3 + 1 : 0, nf - iter0));     // iterations the sequential loop executes
                sh.iter = iter0 + done;
                if (sh.iter >= nf) sh.stop = 1;
            }
        }
        FM_TICK(4);
        __syncthreads();
        if (sh.stop) break;
    }
    if (lane == 0 && scored) atomicAdd(&inf[2], scored);
    if (!sh.have) {
        if (tid == 0) inf[1] = sh.iter;
        return;
    }
    if (lmeds) {
        // inlier threshold of LMedS: sigma = 2.5 * 1.4826 * (1 + 5 / (n - 7)) * sqrt(minMedian), at least 0.001
        if (tid == 0) {
            double sigma = 2.5 * 1.4826 * (1. + 5. / (n - FM_MODEL_POINTS)) * sqrt((double)__int_as_float(sh.min_median));
            sigma = fmax(sigma, 0.001);
            sh.lmeds_t2 = __double2float_rn(sigma * sigma);
        }
        __syncthreads();
        t2 = sh.lmeds_t2;
        tlo = (double)t2 * (1. - 1e-9);
        thi = (double)t2 * (1. + 1e-6);
    }

    // ---- status mask of the winning matrix
    double F[9];
#pragma unroll
    for (int i = 0; i < 9; i++) F[i] = sh.best[i];
    // ---- normalised 8-point on the inliers (run8Point).  Sums run in a fixed tree order, not OpenCV's sequential one.
    double acc[4] = {0, 0, 0, 0};
    int mine = 0;
    for (int i = tid; i < n; i += FM_THREADS) {
        const float2 a = P1[i], b = P2[i];
        const bool in = is_inlier(F, sh.best, a, b, t2, tlo, thi);
        st[i] = in;
        if (in) { acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; mine++; }
    }
    {
        double v[5] = {acc[0], acc[1], acc[2], acc[3], (double)mine};
        block_sum<5, FM_THREADS / 32>(sh, v);
    }
    const int k = (int)sh.red[0][4];
    const double tk = 1. / k;
    const double c1x = sh.red[0][0] * tk, c1y = sh.red[0][1] * tk, c2x = sh.red[0][2] * tk, c2y = sh.red[0][3] * tk;
    if (lmeds && k < FM_MODEL_POINTS) {                  // LMeDSPointSetRegistrator: result = count >= modelPoints
        for (int i = tid; i < n; i += FM_THREADS) st[i] = 0;
        if (tid == 0) inf[1] = sh.iter;
        return;
    }
    if (tid == 0) { inf[0] = k; inf[1] = sh.iter; }
    if (k < 8) return;                                   // 7 inliers: OpenCV's FM_8POINT call degenerates to the 7-point solver
    {
        double v[2] = {0, 0};
        for (int i = tid; i < n; i += FM_THREADS) {
            if (st[i]) {
                const float2 a = P1[i], b = P2[i];
                const double dx1 = a.x - c1x, dy1 = a.y - c1y, dx2 = b.x - c2x, dy2 = b.y - c2y;
                v[0] += sqrt(dx1 * dx1 + dy1 * dy1);
                v[1] += sqrt(dx2 * dx2 + dy2 * dy2);
            }
        }
        block_sum<2, FM_THREADS / 32>(sh, v);
    }
    double s1 = sh.red[0][0] * tk, s2 = sh.red[0][1] * tk;
    if (s1 < (double)FLT_EPSILON || s2 < (double)FLT_EPSILON) return;
    s1 = sqrt(2.) / s1; s2 = sqrt(2.) / s2;
    {
        double v[45];
#pragma unroll
        for (int i = 0; i < 45; i++) v[i] = 0;
        for (int i = tid; i < n; i += FM_THREADS) {
            if (st[i]) {
                const float2 a = P1[i], b = P2[i];
                const double x1 = (a.x - c1x) * s1, y1 = (a.y - c1y) * s1, x2 = (b.x - c2x) * s2, y2 = (b.y - c2y) * s2;
                const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1};
                int e = 0;
#pragma unroll
                for (int p = 0; p < 9; p++)
#pragma unroll
                    for (int q = p; q < 9; q++) v[e++] += r[p] * r[q];
            }
        }
        block_sum<45, FM_THREADS / 32>(sh, v);
    }
    if (tid < 45) {
        int p = 0, e = tid;
        while (e >= 9 - p) { e -= 9 - p; p++; }
        const int q = p + e;
        const double x = sh.red[0][tid];
        sh.A[p * 9 + q] = x;
        sh.A[q * 9 + p] = x;
    }
    __syncthreads();
    FM_TICK(5);
    if (eight_ws) {
        // large batches: the eigen-decomposition is one warp's work; left to k_fm_eight_point (a warp per pair, every pair of the
        // batch at once) it does not hold this CTA's slot while other pairs wait for one
        double* ws = eight_ws + (size_t)pair * FM_EIGHT_WS;
        if (tid < 81) ws[tid] = sh.A[tid];
        if (tid == 0) { ws[81] = c1x; ws[82] = c1y; ws[83] = s1; ws[84] = c2x; ws[85] = c2y; ws[86] = s2; ws[87] = 1.; }
        return;
    }
    if (warp == 0) {
        jacobi9_warp(sh, lane);
        FM_TICK(6);
        if (lane == 0) eight_point_finish(sh, c1x, c1y, s1, c2x, c2y, s2, Fo, inf);
    }
}

// The end of the 8-point step for the pairs whose normal equations k_fm_ransac left in the workspace: one warp per pair.
struct FmEightShared {
    double A[81], V[81];
    double rot_cs[4], rot_sn[4];
    int rot_p[4], rot_q[4];
};
constexpr int FM_EIGHT_WARPS = 2;
__global__ void __launch_bounds__(FM_EIGHT_WARPS * 32)
k_fm_eight_point(const double* __restrict__ eight_ws, int npairs, double* __restrict__ Fout, int32_t* __restrict__ info)
{
    __shared__ FmEightShared s_sh[FM_EIGHT_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = blockIdx.x * FM_EIGHT_WARPS + warp;
    if (pair >= npairs) return;
    const double* ws = eight_ws + (size_t)pair * FM_EIGHT_WS;
    if (ws[87] == 0.) return;
    FmEightShared& sh = s_sh[warp];
    for (int i = lane; i < 81; i += 32) sh.A[i] = ws[i];
    __syncwarp();
    jacobi9_warp(sh, lane);
    if (lane == 0) eight_point_finish(sh, ws[81], ws[82], ws[83], ws[84], ws[85], ws[86], Fout + (size_t)pair * 9, info + (size_t)pair * 4);
}

// Correspondences from the device-resident keypoints and match lists of the sequence mode.  Pair p = f*back + (j-1) is
// (frame f, its j-th predecessor), j = 1..back; predecessors before the batch come from the history [nhist][cap]
// (entry 0 = the frame just before the batch); pairs without a predecessor get count 0.
__global__ void k_fm_gather(const orbx_keypoint* __restrict__ kps, const orbx_keypoint* __restrict__ hist_kps, int nhist, int back, int cap,
                            const orbx_dmatch* __restrict__ good, const long long* __restrict__ ngood,
                            float2* __restrict__ pts1, float2* __restrict__ pts2, int32_t* __restrict__ counts)
{
    const int p = blockIdx.y, f = p / back, j = p - f * back + 1;
    const orbx_keypoint* kq = kps + (size_t)f * cap;
    const orbx_keypoint* kt = nullptr;
    if (f - j >= 0) kt = kps + (size_t)(f - j) * cap;
    else if (hist_kps && j - f - 1 < nhist) kt = hist_kps + (size_t)(j - f - 1) * cap;
    const int n = kt ? (int)min((long long)cap, ngood[p]) : 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[p] = n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const orbx_dmatch m = good[(size_t)p * cap + i];
        const orbx_keypoint a = kq[m.query_idx], b = kt[m.train_idx];
        pts1[(size_t)p * cap + i] = make_float2(a.x, a.y);
        pts2[(size_t)p * cap + i] = make_float2(b.x, b.y);
    }
}

}  // namespace
}  // namespace orbx

using namespace orbx;

struct fmx_context {
    int device;
    cudaStream_t own_stream, stream;
    float2* d_p1; size_t p1_bytes;
    float2* d_p2; size_t p2_bytes;
    int32_t* d_counts; size_t counts_bytes;
    uint8_t* d_status; size_t status_bytes;
    double* d_F; size_t F_bytes;
    int32_t* d_info; size_t info_bytes;
    double* d_eight; size_t eight_bytes;     // normal equations of the pairs of a large batch (k_fm_eight_point)
    int32_t* h_info; size_t h_info_n;
    float* h_pts; size_t h_pts_n;            // pinned staging of fmx_compute_fundamental (grow-only)
    size_t smem_optin;
    int sm_count;
};

template <typename T>
static int fm_grow(T** p, size_t* have, size_t want)
{
    if (*have >= want && *p) return ORBX_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *have = 0; }
    const size_t bytes = align_up(want + want / 4, 256);
    cudaError_t e = cudaMalloc((void**)p, bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return ORBX_E_ALLOC; }
    *have = bytes;
    return ORBX_OK;
}

extern "C" int fmx_create(fmx_handle* out, int device)
{
    ORBX_REQUIRE(out != nullptr, "fmx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { set_error("fmx_create: no CUDA device (%s); liborbx has no CPU fallback", cudaGetErrorString(e)); return ORBX_E_CUDA; }
    ORBX_REQUIRE(device >= 0 && device < ndev, "fmx_create: device %d out of range [0,%d)", device, ndev);
    ORBX_CUDA(cudaSetDevice(device));
    fmx_context* h = new fmx_context();
    memset(h, 0, sizeof(*h));
    h->device = device;
    ORBX_CUDA_OR(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking), fmx_destroy(h));
    h->stream = h->own_stream;
    int optin = 0;
    ORBX_CUDA_OR(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device), fmx_destroy(h));
    h->smem_optin = (size_t)optin;
    ORBX_CUDA_OR(cudaFuncSetAttribute(k_fm_ransac<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin), fmx_destroy(h));
    ORBX_CUDA_OR(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device), fmx_destroy(h));
    *out = h;
    return ORBX_OK;
}

extern "C" int fmx_destroy(fmx_handle h)
{
    if (!h) return ORBX_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_p1); cudaFree(h->d_p2); cudaFree(h->d_counts); cudaFree(h->d_status); cudaFree(h->d_F); cudaFree(h->d_info); cudaFree(h->d_eight);
    if (h->h_info) cudaFreeHost(h->h_info);
    if (h->h_pts) cudaFreeHost(h->h_pts);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return ORBX_OK;
}

extern "C" int fmx_set_stream(fmx_handle h, void* cuda_stream)
{
    ORBX_REQUIRE(h != nullptr, "fmx_set_stream: NULL handle");
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return ORBX_OK;
}

extern "C" int fmx_get_stream(fmx_handle h, void** cuda_stream)
{
    ORBX_REQUIRE(h != nullptr && cuda_stream != nullptr, "fmx_get_stream: NULL argument");
    *cuda_stream = (void*)h->stream;
    return ORBX_OK;
}

extern "C" int fmx_synchronize(fmx_handle h)
{
    ORBX_REQUIRE(h != nullptr, "fmx_synchronize: NULL handle");
    ORBX_CUDA(cudaSetDevice(h->device));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

// max_count: the largest counts[] value if the caller knows it (sizes the shared-memory copy of the points), else `cap`
static int fm_launch(fmx_handle h, const float* d_pts1, const float* d_pts2, const int32_t* d_counts, int npairs, int cap, int max_count,
                     double max_distance, double confidence, uint8_t* d_status, double* d_F, int32_t* d_info)
{
    if (npairs == 0) return ORBX_OK;
    const bool in_smem = max_count <= FM_SMEM_POINTS && sizeof(FmShared) + (size_t)max_count * 16 <= h->smem_optin;
    const size_t smem = sizeof(FmShared) + (in_smem ? (size_t)max_count * 16 : 0);
    const float2* p1 = (const float2*)d_pts1;
    const float2* p2 = (const float2*)d_pts2;
    // 256-thread CTAs (8 warps score one pair's candidates), two per SM.  Measured against 128 x 4, 192 x 3 and 256 x 3 (80
    // registers) at 64 / 296 / 320 / 640 pairs: the batch ends with its slowest pairs, and those finish sooner with more warps
    // and all 128 registers each (profiles/README.md).
    // batches of more pairs than CTA slots (2 per SM): the last step of every pair, one warp's eigen-decomposition, runs in a
    // second kernel for all pairs at once instead of holding a slot
    double* ws = nullptr;
    if (npairs > 2 * h->sm_count) {
        const int rc = fm_grow(&h->d_eight, &h->eight_bytes, (size_t)npairs * FM_EIGHT_WS * sizeof(double));
        if (rc) return rc;
        ws = h->d_eight;
    }
#define FM_LAUNCH(SM) k_fm_ransac<SM, 256><<<npairs, 256, smem, h->stream>>>(p1, p2, d_counts, cap, max_distance, confidence, 1000, d_status, d_F, d_info, ws)
    if (in_smem) FM_LAUNCH(true);
    else FM_LAUNCH(false);
#undef FM_LAUNCH
    ORBX_CUDA(cudaGetLastError());
    if (ws) {
        k_fm_eight_point<<<(npairs + FM_EIGHT_WARPS - 1) / FM_EIGHT_WARPS, FM_EIGHT_WARPS * 32, 0, h->stream>>>(ws, npairs, d_F, d_info);
        ORBX_CUDA(cudaGetLastError());
    }
    return ORBX_OK;
}

extern "C" int fmx_fundamental_batch_dev(fmx_handle h, const float* d_pts1, const float* d_pts2, const int32_t* d_counts, int npairs, int cap,
                                         double max_distance, double confidence, uint8_t* d_status, double* d_F, int32_t* d_info)
{
    ORBX_REQUIRE(h != nullptr, "fmx_fundamental_batch_dev: NULL handle");
    ORBX_REQUIRE(npairs >= 0 && cap >= 1, "fmx_fundamental_batch_dev: npairs %d / cap %d out of range", npairs, cap);
    ORBX_REQUIRE(d_pts1 && d_pts2 && d_counts && d_status && d_F && d_info, "fmx_fundamental_batch_dev: NULL pointer");
    ORBX_CUDA(cudaSetDevice(h->device));
    return fm_launch(h, d_pts1, d_pts2, d_counts, npairs, cap, cap, max_distance, confidence, d_status, d_F, d_info);
}

extern "C" int fmx_fundamental_batch(fmx_handle h, const float* pts1, const float* pts2, const int32_t* counts, int npairs, int cap,
                                     double max_distance, double confidence, uint8_t* status, double* F, int32_t* ninliers)
{
    ORBX_REQUIRE(h != nullptr, "fmx_fundamental_batch: NULL handle");
    ORBX_REQUIRE(npairs >= 0 && cap >= 1, "fmx_fundamental_batch: npairs %d / cap %d out of range", npairs, cap);
    if (npairs == 0) return ORBX_OK;
    ORBX_REQUIRE(pts1 && pts2 && counts && status && F && ninliers, "fmx_fundamental_batch: NULL pointer");
    int max_count = 0;
    for (int i = 0; i < npairs; i++) {
        ORBX_REQUIRE(counts[i] >= 0 && counts[i] <= cap, "fmx_fundamental_batch: counts[%d] = %d outside [0, cap = %d]", i, counts[i], cap);
        max_count = std::max(max_count, counts[i]);
    }
    ORBX_CUDA(cudaSetDevice(h->device));
    const size_t np = (size_t)npairs, pts_bytes = np * cap * sizeof(float2);
    int rc = fm_grow(&h->d_p1, &h->p1_bytes, pts_bytes);
    if (!rc) rc = fm_grow(&h->d_p2, &h->p2_bytes, pts_bytes);
    if (!rc) rc = fm_grow(&h->d_counts, &h->counts_bytes, np * sizeof(int32_t));
    if (!rc) rc = fm_grow(&h->d_status, &h->status_bytes, np * cap);
    if (!rc) rc = fm_grow(&h->d_F, &h->F_bytes, np * 9 * sizeof(double));
    if (!rc) rc = fm_grow(&h->d_info, &h->info_bytes, np * 4 * sizeof(int32_t));
    if (rc) return rc;
    if (h->h_info_n < np * 4) {
        if (h->h_info) cudaFreeHost(h->h_info);
        h->h_info = nullptr; h->h_info_n = 0;
        ORBX_CUDA(cudaMallocHost((void**)&h->h_info, np * 4 * sizeof(int32_t) * 2));
        h->h_info_n = np * 8;
    }
    ORBX_CUDA(cudaMemcpyAsync(h->d_p1, pts1, pts_bytes, cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(h->d_p2, pts2, pts_bytes, cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(h->d_counts, counts, np * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    rc = fm_launch(h, (const float*)h->d_p1, (const float*)h->d_p2, h->d_counts, npairs, cap, std::max(max_count, 1), max_distance, confidence,
                   h->d_status, h->d_F, h->d_info);
    if (rc) return rc;
    ORBX_CUDA(cudaMemcpyAsync(status, h->d_status, np * cap, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(F, h->d_F, np * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(h->h_info, h->d_info, np * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < npairs; i++) ninliers[i] = h->h_info[4 * i];
    return ORBX_OK;
}

extern "C" int fmx_last_info(fmx_handle h, int npairs, int32_t* info)
{
    ORBX_REQUIRE(h != nullptr && info != nullptr, "fmx_last_info: NULL argument");
    ORBX_REQUIRE(npairs >= 0 && (size_t)npairs * 4 <= h->h_info_n, "fmx_last_info: npairs %d exceeds the last batch", npairs);
    memcpy(info, h->h_info, (size_t)npairs * 4 * sizeof(int32_t));
    return ORBX_OK;
}

extern "C" int fmx_compute_fundamental(fmx_handle h, const orbx_keypoint* kps1, int n1, const orbx_keypoint* kps2, int n2,
                                       const orbx_dmatch* matches, int nm, double max_distance, double confidence, uint8_t* status, double* F,
                                       int32_t* ninliers)
{
    ORBX_REQUIRE(h != nullptr, "fmx_compute_fundamental: NULL handle");
    ORBX_REQUIRE(nm >= 0 && n1 >= 0 && n2 >= 0, "fmx_compute_fundamental: negative size");
    ORBX_REQUIRE(F && ninliers && (nm == 0 || (kps1 && kps2 && matches && status)), "fmx_compute_fundamental: NULL pointer");
    *ninliers = 0;
    for (int i = 0; i < 9; i++) F[i] = 0.;
    if (nm == 0) return ORBX_OK;
    if (h->h_pts_n < (size_t)nm * 4) {
        if (h->h_pts) cudaFreeHost(h->h_pts);
        h->h_pts = nullptr; h->h_pts_n = 0;
        const size_t want = (size_t)nm * 4 + (size_t)nm;
        ORBX_CUDA(cudaMallocHost((void**)&h->h_pts, want * sizeof(float)));
        h->h_pts_n = want;
    }
    float* a = h->h_pts;
    float* b = h->h_pts + 2 * (size_t)nm;
    for (int i = 0; i < nm; i++) {
        const int q = matches[i].query_idx, t = matches[i].train_idx;
        if (q < 0 || q >= n1 || t < 0 || t >= n2) {
            set_error("fmx_compute_fundamental: match %d indexes (%d, %d) outside (%d, %d) keypoints", i, q, t, n1, n2);
            return ORBX_E_INVALID;
        }
        a[2 * i] = kps1[q].x; a[2 * i + 1] = kps1[q].y;
        b[2 * i] = kps2[t].x; b[2 * i + 1] = kps2[t].y;
    }
    const int32_t count = nm;
    return fmx_fundamental_batch(h, a, b, &count, 1, nm, max_distance, confidence, status, F, ninliers);
}

extern "C" int fmx_filter_back_dev(fmx_handle h, const orbx_keypoint* d_kps, int nframes, int cap, int back, const orbx_keypoint* d_hist_kps,
                                   int nhist, const orbx_dmatch* d_good, const int64_t* d_ngood, double max_distance, double confidence,
                                   uint8_t* d_status, double* d_F, int32_t* d_info)
{
    ORBX_REQUIRE(h != nullptr, "fmx_filter_back_dev: NULL handle");
    ORBX_REQUIRE(nframes >= 0 && cap >= 1 && back >= 1 && nhist >= 0, "fmx_filter_back_dev: nframes %d / cap %d / back %d / nhist %d out of range",
                 nframes, cap, back, nhist);
    if (nframes == 0) return ORBX_OK;
    ORBX_REQUIRE(d_kps && d_good && d_ngood && d_status && d_F && d_info, "fmx_filter_back_dev: NULL pointer");
    ORBX_CUDA(cudaSetDevice(h->device));
    const int npairs = nframes * back;
    const size_t np = (size_t)npairs, pts_bytes = np * cap * sizeof(float2);
    int rc = fm_grow(&h->d_p1, &h->p1_bytes, pts_bytes);
    if (!rc) rc = fm_grow(&h->d_p2, &h->p2_bytes, pts_bytes);
    if (!rc) rc = fm_grow(&h->d_counts, &h->counts_bytes, np * sizeof(int32_t));
    if (rc) return rc;
    k_fm_gather<<<dim3(div_up(cap, 256 * 2), npairs), 256, 0, h->stream>>>(d_kps, d_hist_kps, nhist, back, cap, d_good, (const long long*)d_ngood,
                                                                          h->d_p1, h->d_p2, h->d_counts);
    ORBX_CUDA(cudaGetLastError());
    return fm_launch(h, (const float*)h->d_p1, (const float*)h->d_p2, h->d_counts, npairs, cap, cap, max_distance, confidence, d_status, d_F,
                     d_info);
}

extern "C" int fmx_filter_consecutive_dev(fmx_handle h, const orbx_keypoint* d_kps, const orbx_keypoint* d_prev_kps, int nframes, int cap,
                                          const orbx_dmatch* d_good, const int64_t* d_ngood, double max_distance, double confidence,
                                          uint8_t* d_status, double* d_F, int32_t* d_info)
{
    return fmx_filter_back_dev(h, d_kps, nframes, cap, 1, d_prev_kps, d_prev_kps ? 1 : 0, d_good, d_ngood, max_distance, confidence, d_status,
                               d_F, d_info);
}

#ifdef FM_PROFILE
extern "C" __attribute__((visibility("default"))) int fmx_debug_profile(unsigned long long* out8, int reset)
{
    ORBX_CUDA(cudaDeviceSynchronize());
    ORBX_CUDA(cudaMemcpyFromSymbol(out8, g_fm_prof, sizeof(g_fm_prof)));
    if (reset) { unsigned long long z[8] = {0}; ORBX_CUDA(cudaMemcpyToSymbol(g_fm_prof, z, sizeof(z))); }
    return ORBX_OK;
}
// start / end %globaltimer (ns) of the first npairs CTAs of the last launch
extern "C" __attribute__((visibility("default"))) int fmx_debug_pair_times(unsigned long long* out, int npairs)
{
    ORBX_REQUIRE(npairs >= 0 && npairs <= FM_PROF_PAIRS, "fmx_debug_pair_times: %d pairs outside [0, %d]", npairs, FM_PROF_PAIRS);
    ORBX_CUDA(cudaDeviceSynchronize());
    ORBX_CUDA(cudaMemcpyFromSymbol(out, g_fm_pair_ns, (size_t)npairs * 2 * sizeof(unsigned long long)));
    return ORBX_OK;
}
#endif
