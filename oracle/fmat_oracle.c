/*
 * fmat_oracle.c -- CPU restatement of the reference's fundamental-matrix outlier filter.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as orb_oracle.c): used by tests/, __graft_entry__.smoke() and bench.py's CPU legs
 * as the checker / reported baseline.  The product library never links or calls it.
 *
 * Path restated: computeFundamentalMatrix (reference src/CameraPoseEstimator.cpp:545-586):
 *     F = findFundamentalMat(inputs1, inputs2, CV_FM_RANSAC, MAX_DISTANCE=3, CONFIDENCE=0.85, status)   (:563, ParamConfig.h:24-25)
 *     F = findFundamentalMat(inliers1, inliers2, CV_FM_8POINT)                                          (:585)
 * The arithmetic lives in OpenCV's calib3d (fundam.cpp, ptsetreg.cpp), not under /root/reference.  As for the ORB path
 * the parity target is the runnable cv2 4.13.0: this file follows its published algorithm and is pinned against
 * cv2.findFundamentalMat outputs committed in tests/golden/fmat_cases.npz (tests/golden/make_golden_fmat.py): inlier
 * masks identical, F within 1e-6 relative (Frobenius, after scaling to F[8] = 1).
 *
 *   - RANSAC over 7-point samples for n >= 15 points (cv::RNG seeded with (uint64)-1, duplicate-free index draws, the
 *     collinearity test on the last drawn point, at most 3 models per sample, symmetric epipolar distance in double
 *     rounded to float, strict ">" best update, adaptive iteration count);
 *   - LMedS (same sampler, fixed iteration count from a 0.45 outlier ratio) for 8 <= n < 15, as OpenCV does;
 *   - normalised 8-point on the inliers, rank 2 enforced.
 * Linear algebra (null space, symmetric eigenvectors) is done here with Householder QR and cyclic Jacobi; OpenCV uses its
 * own Jacobi SVD.  The 2-D null space of a rank-7 system is unique, so the candidate matrices agree to rounding.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---- cv::RNG (multiply-with-carry), modules/core/include/opencv2/core/operations.hpp ---- */
typedef struct { uint64_t s; } fm_rng;
static unsigned rng_next(fm_rng* r) {
    r->s = (uint64_t)(unsigned)r->s * 4164903690U + (unsigned)(r->s >> 32);
    return (unsigned)r->s;
}
static int rng_uniform(fm_rng* r, int a, int b) { return a == b ? a : (int)(rng_next(r) % (unsigned)(b - a) + a); }

/* last point of the subset on a line through two earlier ones (or too close to one) */
static int have_collinear(const float* p /* count x 2 */, int count) {
    const int i = count - 1;
    for (int j = 0; j < i; j++) {
        const double dx1 = (double)p[2 * j] - p[2 * i], dy1 = (double)p[2 * j + 1] - p[2 * i + 1];
        for (int k = 0; k < j; k++) {
            const double dx2 = (double)p[2 * k] - p[2 * i], dy2 = (double)p[2 * k + 1] - p[2 * i + 1];
            if (fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return 1;
        }
    }
    return 0;
}

static int get_subset(const float* m1, const float* m2, int count, fm_rng* rng, int max_attempts, float* s1, float* s2) {
    int idx[7], i = 0, iters = 0;
    for (; iters < max_attempts; iters++) {
        for (i = 0; i < 7 && iters < max_attempts;) {
            int idx_i, j;
            for (;;) {
                idx_i = idx[i] = rng_uniform(rng, 0, count);
                for (j = 0; j < i; j++)
                    if (idx_i == idx[j]) break;
                if (j == i) break;
            }
            s1[2 * i] = m1[2 * idx_i]; s1[2 * i + 1] = m1[2 * idx_i + 1];
            s2[2 * i] = m2[2 * idx_i]; s2[2 * i + 1] = m2[2 * idx_i + 1];
            i++;
        }
        if (i == 7 && (have_collinear(s1, 7) || have_collinear(s2, 7))) continue;
        break;
    }
    return i == 7 && iters < max_attempts;
}

/* cv::solveCubic (modules/core/src/mathfuncs.cpp); returns the number of real roots, in OpenCV's order */
ORC_API int orc_solve_cubic(const double* c, double* x) {
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
    double d = Qcubed - R * R;
    if (d > 0) {
        const double theta = acos(R / sqrt(Qcubed)), sqrtQ = sqrt(Q);
        const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
        x[0] = t0 * cos(t1) - t2;
        x[1] = t0 * cos(t1 + (2. * M_PI / 3)) - t2;
        x[2] = t0 * cos(t1 + (4. * M_PI / 3)) - t2;
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

/* centre and isotropic scale (mean distance sqrt 2) of a point set, as run7Point / run8Point compute them */
static int normalisation(const float* m, int count, double* cx, double* cy, double* scale) {
    double sx = 0, sy = 0, sc = 0;
    for (int i = 0; i < count; i++) { sx += m[2 * i]; sy += m[2 * i + 1]; }
    const double t = 1. / count;
    sx *= t; sy *= t;
    for (int i = 0; i < count; i++) {
        const double dx = m[2 * i] - sx, dy = m[2 * i + 1] - sy;
        sc += sqrt(dx * dx + dy * dy);
    }
    sc *= t;
    *cx = sx; *cy = sy;
    if (sc < FLT_EPSILON) return 0;
    *scale = sqrt(2.) / sc;
    return 1;
}

/* orthonormal basis (n1, n2) of the null space of the 7x9 matrix a (row-major): Householder QR of its transpose */
static void null_space_7x9(const double* a, double* n1, double* n2) {
    double m[9][7], v[7][9];
    for (int r = 0; r < 7; r++)
        for (int c = 0; c < 9; c++) m[c][r] = a[r * 9 + c];
    for (int k = 0; k < 7; k++) {
        double norm = 0;
        for (int i = k; i < 9; i++) norm += m[i][k] * m[i][k];
        norm = sqrt(norm);
        for (int i = 0; i < 9; i++) v[k][i] = 0;
        if (norm == 0) continue;
        const double alpha = m[k][k] > 0 ? -norm : norm;
        double vn = 0;
        for (int i = k; i < 9; i++) { v[k][i] = m[i][k]; }
        v[k][k] -= alpha;
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
    double* out[2] = {n1, n2};
    for (int j = 0; j < 2; j++) {
        double q[9] = {0};
        q[7 + j] = 1;
        for (int k = 6; k >= 0; k--) {
            double dot = 0;
            for (int i = k; i < 9; i++) dot += v[k][i] * q[i];
            for (int i = k; i < 9; i++) q[i] -= 2 * dot * v[k][i];
        }
        memcpy(out[j], q, sizeof q);
    }
}

static void denormalise(double* F, double c1x, double c1y, double s1, double c2x, double c2y, double s2) {
    /* F <- T2' * F * T1, then F *= 1/F[8] when |F[8]| > FLT_EPSILON */
    const double T1[9] = {s1, 0, -s1 * c1x, 0, s1, -s1 * c1y, 0, 0, 1}, T2[9] = {s2, 0, -s2 * c2x, 0, s2, -s2 * c2y, 0, 0, 1};
    double G[9], H[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += T2[k * 3 + i] * F[k * 3 + j];
            G[i * 3 + j] = s;
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += G[i * 3 + k] * T1[k * 3 + j];
            H[i * 3 + j] = s;
        }
    if (fabs(H[8]) > FLT_EPSILON) {
        const double t = 1. / H[8];
        for (int i = 0; i < 9; i++) H[i] *= t;
    }
    memcpy(F, H, sizeof H);
}

/* run7Point: up to 3 candidate matrices (9 doubles each) from 7 correspondences; returns their number */
ORC_API int orc_fm_7point(const float* m1, const float* m2, double* models /* 27 */) {
    double c1x, c1y, s1, c2x, c2y, s2, a[63], f1[9], f2[9], c[4], r[3] = {0, 0, 0};
    if (!normalisation(m1, 7, &c1x, &c1y, &s1) || !normalisation(m2, 7, &c2x, &c2y, &s2)) return 0;
    for (int i = 0; i < 7; i++) {
        const double x0 = (m1[2 * i] - c1x) * s1, y0 = (m1[2 * i + 1] - c1y) * s1;
        const double x1 = (m2[2 * i] - c2x) * s2, y1 = (m2[2 * i + 1] - c2y) * s2;
        double* row = a + 9 * i;
        row[0] = x1 * x0; row[1] = x1 * y0; row[2] = x1;
        row[3] = y1 * x0; row[4] = y1 * y0; row[5] = y1;
        row[6] = x0; row[7] = y0; row[8] = 1;
    }
    null_space_7x9(a, f1, f2);
    /* F = lambda*f1 + (1-lambda)*f2; det F = 0 is a cubic in lambda */
    for (int i = 0; i < 9; i++) f1[i] -= f2[i];
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
    const int n = orc_solve_cubic(c, r);
    if (n < 1 || n > 3) return n < 0 ? 0 : n;
    for (int k = 0; k < n; k++) {
        double* F = models + 9 * k;
        double lambda = r[k], mu = 1.;
        const double s = f1[8] * r[k] + f2[8];
        if (fabs(s) > DBL_EPSILON) { mu = 1. / s; lambda *= mu; F[8] = 1.; }
        else F[8] = 0.;
        for (int i = 0; i < 8; i++) F[i] = f1[i] * lambda + f2[i] * mu;
        denormalise(F, c1x, c1y, s1, c2x, c2y, s2);
    }
    return n;
}

/* cyclic Jacobi for a symmetric n x n matrix (n <= 9): eigenvalues w, eigenvectors as the ROWS of v, sorted descending */
static void jacobi_eigen(double* A, int n, double* w, double* v) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) v[i * n + j] = i == j;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0, diag = 0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                if (i != j) off += A[i * n + j] * A[i * n + j];
                else diag += A[i * n + j] * A[i * n + j];
            }
        if (off <= 1e-60 * diag || off == 0) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = A[p * n + q];
                if (apq == 0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
                const double t = (theta >= 0 ? 1. : -1.) / (fabs(theta) + sqrt(theta * theta + 1));
                const double cs = 1. / sqrt(t * t + 1), sn = t * cs;
                for (int k = 0; k < n; k++) {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = cs * akp - sn * akq;
                    A[k * n + q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = cs * apk - sn * aqk;
                    A[q * n + k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double vpk = v[p * n + k], vqk = v[q * n + k];
                    v[p * n + k] = cs * vpk - sn * vqk;
                    v[q * n + k] = sn * vpk + cs * vqk;
                }
            }
    }
    for (int i = 0; i < n; i++) w[i] = A[i * n + i];
    for (int i = 0; i < n - 1; i++) {
        int m = i;
        for (int j = i + 1; j < n; j++)
            if (w[j] > w[m]) m = j;
        if (m != i) {
            double t = w[i]; w[i] = w[m]; w[m] = t;
            for (int k = 0; k < n; k++) { t = v[i * n + k]; v[i * n + k] = v[m * n + k]; v[m * n + k] = t; }
        }
    }
}

/* run8Point: normalised 8-point over count >= 8 correspondences, rank 2 enforced; returns 1 or 0 (degenerate) */
ORC_API int orc_fm_8point(const float* m1, const float* m2, int count, double* F) {
    double c1x, c1y, s1, c2x, c2y, s2, A[81], W[9], V[81];
    if (count < 8) return 0;
    if (!normalisation(m1, count, &c1x, &c1y, &s1) || !normalisation(m2, count, &c2x, &c2y, &s2)) return 0;
    memset(A, 0, sizeof A);
    for (int i = 0; i < count; i++) {
        const double x1 = (m1[2 * i] - c1x) * s1, y1 = (m1[2 * i + 1] - c1y) * s1;
        const double x2 = (m2[2 * i] - c2x) * s2, y2 = (m2[2 * i + 1] - c2y) * s2;
        const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1};
        for (int j = 0; j < 9; j++)
            for (int k = 0; k < 9; k++) A[j * 9 + k] += r[j] * r[k];
    }
    jacobi_eigen(A, 9, W, V);
    int i;
    for (i = 0; i < 9; i++)
        if (fabs(W[i]) < DBL_EPSILON) break;
    if (i < 8) return 0;
    double F0[9];
    memcpy(F0, V + 72, sizeof F0);          /* eigenvector of the smallest eigenvalue */
    /* rank 2: remove the component along the smallest right singular vector, F0 - (F0 v) v' */
    double G[9], w3[3], V3[9];
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += F0[k * 3 + a] * F0[k * 3 + b];
            G[a * 3 + b] = s;
        }
    jacobi_eigen(G, 3, w3, V3);
    const double* vs = V3 + 6;
    for (int a = 0; a < 3; a++) {
        const double fv = F0[a * 3] * vs[0] + F0[a * 3 + 1] * vs[1] + F0[a * 3 + 2] * vs[2];
        for (int b = 0; b < 3; b++) F0[a * 3 + b] -= fv * vs[b];
    }
    denormalise(F0, c1x, c1y, s1, c2x, c2y, s2);
    memcpy(F, F0, sizeof F0);
    return 1;
}

/* FMEstimatorCallback::computeError: max of the two squared point-to-epipolar-line distances, double rounded to float */
ORC_API void orc_fm_errors(const float* m1, const float* m2, int count, const double* F, float* err) {
    for (int i = 0; i < count; i++) {
        const double x1 = m1[2 * i], y1 = m1[2 * i + 1], x2 = m2[2 * i], y2 = m2[2 * i + 1];
        double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
        const double s2 = 1. / (a * a + b * b), d2 = x2 * a + y2 * b + c;
        a = F[0] * x2 + F[3] * y2 + F[6]; b = F[1] * x2 + F[4] * y2 + F[7]; c = F[2] * x2 + F[5] * y2 + F[8];
        const double s1 = 1. / (a * a + b * b), d1 = x1 * a + y1 * b + c;
        const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
        err[i] = (float)(e1 < e2 ? e2 : e1);          /* std::max(d1*d1*s1, d2*d2*s2) */
    }
}

static int update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = p < 0 ? 0 : p > 1 ? 1 : p;
    ep = ep < 0 ? 0 : ep > 1 ? 1 : ep;
    double num = 1 - p > DBL_MIN ? 1 - p : DBL_MIN;
    double denom = 1 - pow(1 - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)lrint(num / denom);
}

static int cmp_float(const void* a, const void* b) {
    const float x = *(const float*)a, y = *(const float*)b;
    return x < y ? -1 : x > y;
}

/*
 * findFundamentalMat(m1, m2, FM_RANSAC, thr, conf, mask) for n >= 8 points (n == 7 and n < 7: see orc_fm_find).
 * Returns 1 and fills mask (n bytes, 0/1) and F (the model RANSAC / LMedS ended with), 0 when no model was found
 * (mask zeroed).  info[0] = iterations run, info[1] = inliers.
 */
ORC_API int orc_fm_ransac(const float* m1, const float* m2, int n, double thr, double conf, int max_iters, uint8_t* mask,
                          double* F, int* info) {
    if (thr <= 0) thr = 3;
    if (conf < DBL_EPSILON || conf > 1 - DBL_EPSILON) conf = 0.99;
    const int lmeds = n < 15;
    fm_rng rng = {(uint64_t)-1};
    float* err = (float*)malloc(sizeof(float) * (size_t)n * 2);
    float* srt = err + n;
    double models[27], best[9];
    float s1[14], s2[14];
    int niters = lmeds ? update_num_iters(conf, 0.45, 7, max_iters) : max_iters, max_good = 0, iter, have = 0;
    if (lmeds && niters < 3) niters = 3;
    double min_median = DBL_MAX;
    const float t2 = (float)(thr * thr);
    memset(mask, 0, (size_t)n);
    for (iter = 0; iter < niters; iter++) {
        if (!get_subset(m1, m2, n, &rng, lmeds ? 1000 : 10000, s1, s2)) break;
        const int nmodels = orc_fm_7point(s1, s2, models);
        for (int k = 0; k < nmodels; k++) {
            orc_fm_errors(m1, m2, n, models + 9 * k, err);
            if (lmeds) {
                memcpy(srt, err, sizeof(float) * (size_t)n);
                qsort(srt, (size_t)n, sizeof(float), cmp_float);
                const double median = srt[n / 2];
                if (median < min_median) { min_median = median; memcpy(best, models + 9 * k, sizeof best); have = 1; }
            } else {
                int good = 0;
                for (int i = 0; i < n; i++) good += err[i] <= t2;
                if (good > (max_good > 6 ? max_good : 6)) {
                    max_good = good;
                    memcpy(best, models + 9 * k, sizeof best);
                    have = 1;
                    niters = update_num_iters(conf, (double)(n - good) / n, 7, niters);
                }
            }
        }
    }
    int ninl = 0;
    if (have) {
        float t = t2;
        if (lmeds) {
            double sigma = 2.5 * 1.4826 * (1 + 5. / (n - 7)) * sqrt(min_median);
            if (sigma < 0.001) sigma = 0.001;
            t = (float)(sigma * sigma);
        }
        orc_fm_errors(m1, m2, n, best, err);
        for (int i = 0; i < n; i++) { mask[i] = err[i] <= t; ninl += mask[i]; }
        if (lmeds && ninl < 7) { have = 0; memset(mask, 0, (size_t)n); ninl = 0; }
        memcpy(F, best, sizeof best);
    }
    if (info) { info[0] = iter; info[1] = ninl; }
    free(err);
    return have;
}

typedef struct { int32_t query_idx, train_idx, img_idx; float distance; } orc_dmatch_t;

/*
 * computeFundamentalMatrix (reference src/CameraPoseEstimator.cpp:545-586): positions of the matched keypoints
 * (pos1[query_idx], pos2[train_idx]; float x,y pairs with the given strides in floats), RANSAC status per match, then the
 * 8-point matrix of the inliers.  Returns the number of inliers (0: fewer than 8 matches or no model; F zeroed).
 */
ORC_API int orc_compute_fundamental(const float* pos1, int stride1, const float* pos2, int stride2, const void* matches, int nm,
                                    double thr, double conf, uint8_t* status, double* F) {
    const orc_dmatch_t* mt = (const orc_dmatch_t*)matches;
    memset(F, 0, 9 * sizeof(double));
    memset(status, 0, (size_t)(nm > 0 ? nm : 0));
    if (nm < 8) return 0;
    float* a = (float*)malloc(sizeof(float) * 4 * (size_t)nm);
    float* b = a + 2 * (size_t)nm;
    for (int i = 0; i < nm; i++) {
        a[2 * i] = pos1[(size_t)mt[i].query_idx * stride1]; a[2 * i + 1] = pos1[(size_t)mt[i].query_idx * stride1 + 1];
        b[2 * i] = pos2[(size_t)mt[i].train_idx * stride2]; b[2 * i + 1] = pos2[(size_t)mt[i].train_idx * stride2 + 1];
    }
    double Fr[9];
    int info[2] = {0, 0}, ninl = 0;
    if (orc_fm_ransac(a, b, nm, thr, conf, 1000, status, Fr, info)) {
        int k = 0;
        for (int i = 0; i < nm; i++)
            if (status[i]) { a[2 * k] = a[2 * i]; a[2 * k + 1] = a[2 * i + 1]; b[2 * k] = b[2 * i]; b[2 * k + 1] = b[2 * i + 1]; k++; }
        ninl = k;
        if (k < 8 || !orc_fm_8point(a, b, k, F)) memset(F, 0, 9 * sizeof(double));
    }
    free(a);
    return ninl;
}
