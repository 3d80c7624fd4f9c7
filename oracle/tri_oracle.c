/*
 * tri_oracle.c -- CPU restatement of the reference's linear two-view triangulation and of the map-point association
 * loops that consume the match lists.  TEST INFRASTRUCTURE ONLY (see the header of orb_oracle.c): it is the checker of
 * the CUDA path in monocular_slam_b200/csrc/triangulate.cu, never part of the product.
 *
 * What it follows (reference file:line):
 *   TriangulateSinglePointFromTwoView     src/CameraPoseEstimator.cpp:86-132
 *       P1 = K1 * Rt1, P2 = K2 * Rt2                                              :91
 *       A (4x4): rows P[0,:] - P[2,:] * x and P[1,:] - P[2,:] * y of both views   :94-111
 *       solveHLS(A, X): cv::SVD(A, MODIFY_A), X = last row of vt                  :114, src/CommonMath.cpp:17-22
 *       dehomogenize: X /= X[3]                                                   :115, src/CommonMath.cpp:24-27
 *       front of both cameras: (Rt1 * X)[2] > 0 && (Rt2 * X)[2] > 0               :119-127
 *   TriangulateMultiplePointsFromTwoView  src/CameraPoseEstimator.cpp:134-152  (count of points in front when asked)
 *   the bootstrap's four-hypothesis test  src/CameraPoseEstimator.cpp:334-349  (first strict maximum of the counts)
 *   association loop                      src/CameraPoseEstimator.cpp:402-455
 *   new-map-point loop                    src/CameraPoseEstimator.cpp:488-512
 *
 * cv::SVD lives in OpenCV (un-vendored).  Its own implementation is the one-sided Jacobi method of Hestenes on the rows of
 * A^T (modules/core/src/lapack.cpp, JacobiSVDImpl_), restated below from that published algorithm: rotate row pairs until
 * every pair is orthogonal to DBL_EPSILON * 10, singular values = row norms, sorted in descending order.  The cv2 build in
 * this container answers cv2.SVDecomp with LAPACK instead, so the pin is numerical, not bit-wise: tests/golden/tri_cases.npz
 * holds cv2.SVDecomp / cv2.triangulatePoints results and the oracle must reproduce X to 1e-9 relative and every
 * front-of-camera flag exactly (tests/test_oracle_tri.py).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

/* One-sided Jacobi SVD of a 4x4 matrix given as At (row i of At = column i of A).  On return row 3 of Vt is the right
 * singular vector of the smallest singular value. */
static void jacobi_svd4(double At[4][4], double W[4], double Vt[4][4])
{
    const double eps = DBL_EPSILON * 10;
    for (int i = 0; i < 4; i++) {
        double sd = 0;
        for (int k = 0; k < 4; k++) { sd += At[i][k] * At[i][k]; Vt[i][k] = i == k ? 1. : 0.; }
        W[i] = sd;
    }
    for (int iter = 0; iter < 30; iter++) {
        int changed = 0;
        for (int i = 0; i < 3; i++)
            for (int j = i + 1; j < 4; j++) {
                double a = W[i], b = W[j], p = 0;
                for (int k = 0; k < 4; k++) p += At[i][k] * At[j][k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < 4; k++) {
                    const double t0 = c * At[i][k] + s * At[j][k], t1 = -s * At[i][k] + c * At[j][k];
                    At[i][k] = t0; At[j][k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = 1;
                for (int k = 0; k < 4; k++) {
                    const double t0 = c * Vt[i][k] + s * Vt[j][k], t1 = -s * Vt[i][k] + c * Vt[j][k];
                    Vt[i][k] = t0; Vt[j][k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < 4; i++) {
        double sd = 0;
        for (int k = 0; k < 4; k++) sd += At[i][k] * At[i][k];
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < 3; i++) {          /* descending singular values */
        int j = i;
        for (int k = i + 1; k < 4; k++) if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            for (int k = 0; k < 4; k++) {
                t = At[i][k]; At[i][k] = At[j][k]; At[j][k] = t;
                t = Vt[i][k]; Vt[i][k] = Vt[j][k]; Vt[j][k] = t;
            }
        }
    }
}

static void mul_k_rt(const double* K, const double* Rt, double* P)   /* 3x3 * 3x4, row-major */
{
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += K[3 * r + k] * Rt[4 * k + c];
            P[4 * r + c] = s;
        }
}

/* TriangulateSinglePointFromTwoView; returns the front-of-both-cameras flag, X receives the dehomogenised point. */
int orc_triangulate_one(const double* p1, const double* p2, const double* Rt1, const double* Rt2, const double* K1, const double* K2, double* X)
{
    double P1[12], P2[12], At[4][4], W[4], Vt[4][4];
    mul_k_rt(K1, Rt1, P1);
    mul_k_rt(K2, Rt2, P2);
    for (int c = 0; c < 4; c++) {            /* At[c][r] = A[r][c] */
        At[c][0] = P1[c] - P1[8 + c] * p1[0];
        At[c][1] = P1[4 + c] - P1[8 + c] * p1[1];
        At[c][2] = P2[c] - P2[8 + c] * p2[0];
        At[c][3] = P2[4 + c] - P2[8 + c] * p2[1];
    }
    jacobi_svd4(At, W, Vt);
    double Xh[4];
    for (int k = 0; k < 4; k++) Xh[k] = Vt[3][k] / Vt[3][3];
    X[0] = Xh[0]; X[1] = Xh[1]; X[2] = Xh[2];
    double z1 = 0, z2 = 0;
    for (int k = 0; k < 4; k++) { z1 += Rt1[8 + k] * Xh[k]; z2 += Rt2[8 + k] * Xh[k]; }
    return z1 > 0 && z2 > 0;
}

/* TriangulateMultiplePointsFromTwoView: pts are n x 2 doubles, X n x 3, front n bytes (may be NULL); returns the count. */
int orc_triangulate(const double* pts1, const double* pts2, int n, const double* Rt1, const double* Rt2, const double* K1,
                    const double* K2, double* X, uint8_t* front)
{
    int count = 0;
    for (int i = 0; i < n; i++) {
        const int f = orc_triangulate_one(pts1 + 2 * i, pts2 + 2 * i, Rt1, Rt2, K1, K2, X + 3 * i);
        if (front) front[i] = (uint8_t)f;
        count += f;
    }
    return count;
}

/* src/CameraPoseEstimator.cpp:334-349: the four [R|t] candidates against the identity camera; index of the first strict
 * maximum of the front counts.  Rts is nhyp x 12; counts receives nhyp entries; X (n x 3) the points of the winner. */
int orc_triangulate_best(const double* pts1, const double* pts2, int n, const double* Rt1, const double* Rts, int nhyp,
                         const double* K1, const double* K2, double* X, int32_t* counts)
{
    int best = -1, max_count = -1;
    double tmp[3];
    for (int h = 0; h < nhyp; h++) {
        int count = 0;
        for (int i = 0; i < n; i++) count += orc_triangulate_one(pts1 + 2 * i, pts2 + 2 * i, Rt1, Rts + 12 * h, K1, K2, tmp);
        counts[h] = count;
        if (max_count < count) { max_count = count; best = h; }
    }
    if (best >= 0 && X)
        for (int i = 0; i < n; i++) orc_triangulate_one(pts1 + 2 * i, pts2 + 2 * i, Rt1, Rts + 12 * best, K1, K2, X + 3 * i);
    return best;
}

typedef struct { int32_t query_idx, train_idx, img_idx; float distance; } orc_dmatch;

/* The association loop of pnpPoseEstimation (src/CameraPoseEstimator.cpp:402-455) for ONE current frame: `back` match
 * lists (list l = the matches against predecessor l, most recent first, already reduced to the RANSAC inliers as the
 * FILTERING_WITH_F block does), list l holding n[l] matches at matches + l * cap, and the predecessors' map-point indices
 * premap + l * cap (-1 = none).  A match associates its query feature with the predecessor's map point when that feature
 * has none yet; first come, first served.  cur_map (ncur entries, all -1 on entry) receives the map point of every current
 * feature; assoc_q / assoc_mp list the associations in the order the reference discovers them.  Returns their number. */
int orc_associate(const orc_dmatch* matches, const int32_t* n, int back, int cap, const int32_t* premap, int ncur,
                  int32_t* cur_map, int32_t* assoc_q, int32_t* assoc_mp)
{
    int count = 0;
    for (int i = 0; i < ncur; i++) cur_map[i] = -1;
    for (int l = 0; l < back; l++) {
        for (int j = 0; j < n[l]; j++) {
            const int q = matches[(size_t)l * cap + j].query_idx, t = matches[(size_t)l * cap + j].train_idx;
            const int mp = premap[(size_t)l * cap + t];
            if (mp != -1 && cur_map[q] == -1) {       /* cur_map[q] != -1  <=>  matched[q] */
                cur_map[q] = mp;
                assoc_q[count] = q;
                assoc_mp[count] = mp;
                count++;
                if (count == ncur) return count;
            }
        }
        if (count == ncur) break;
    }
    return count;
}

/* The new-map-point loop (src/CameraPoseEstimator.cpp:488-512): walk the cached lists in the same order; a match whose two
 * features are both without a map point is triangulated and registered, which gives both features the new point's index --
 * so a later match that shares either feature is skipped.  accept (back x cap bytes) marks the matches the loop takes;
 * cur_map / premap are updated in place with ids counted from next_id.  Returns the number of new points. */
int orc_select_new(const orc_dmatch* matches, const int32_t* n, int back, int cap, int32_t* premap, int32_t* cur_map,
                   int32_t next_id, uint8_t* accept)
{
    int count = 0;
    for (int l = 0; l < back; l++)
        for (int j = 0; j < n[l]; j++) {
            const int q = matches[(size_t)l * cap + j].query_idx, t = matches[(size_t)l * cap + j].train_idx;
            accept[(size_t)l * cap + j] = 0;
            if (premap[(size_t)l * cap + t] == -1 && cur_map[q] == -1) {
                premap[(size_t)l * cap + t] = next_id + count;
                cur_map[q] = next_id + count;
                accept[(size_t)l * cap + j] = 1;
                count++;
            }
        }
    return count;
}
