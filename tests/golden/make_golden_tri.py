#!/usr/bin/env python
"""Golden vectors for the linear two-view triangulation (tests/golden/tri_cases.npz).

TriangulateSinglePointFromTwoView (reference src/CameraPoseEstimator.cpp:86-132) builds a 4x4 system from the two camera
matrices and the matched pixel positions and takes its null direction from cv::SVD (src/CommonMath.cpp:17-22).  OpenCV is
not vendored in the reference; as for the other stages, the runnable cv2 4.13.0 of this container is the parity target.
For every seeded two-view scene below this script restates the reference's routine call by call with cv2.SVDecomp
(the Python binding of the same cv::SVD::compute) and stores X, the front-of-both-cameras flags and their count; it also
stores cv2.triangulatePoints' answer for the same cameras (the same DLT system up to row signs) as a second opinion.

Run (build container only; needs cv2):  python tests/golden/make_golden_tri.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def rot(ax, ay, az):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def scene(seed, n, noise, outliers, size=(1920, 1080), general_first=False):
    """n matched positions of a synthetic two-view pair.  Positions are float32 keypoint coordinates widened to double, as
    the reference stores them (src/FeatureExtractor.cpp:20-22: Point2d from KeyPoint.pt)."""
    r = np.random.default_rng(seed)
    w, h = size
    K1 = np.array([[900.0 + r.uniform(-50, 50), 0, w / 2 + r.uniform(-20, 20)], [0, 900.0 + r.uniform(-50, 50), h / 2 + r.uniform(-20, 20)], [0, 0, 1]])
    K2 = K1.copy() if seed % 2 else np.array([[870.0, 0, w / 2 - 7], [0, 880.0, h / 2 + 4], [0, 0, 1]])
    if general_first:
        Rt1 = np.c_[rot(0.02, -0.03, 0.01), np.array([0.1, -0.05, 0.02])]
    else:
        Rt1 = np.c_[np.eye(3), np.zeros(3)]                       # the bootstrap's identity camera (:333)
    R2 = rot(*r.uniform(-0.08, 0.08, 3))
    t2 = np.array([r.uniform(0.2, 0.6) * r.choice([-1, 1]), r.uniform(-0.1, 0.1), r.uniform(-0.15, 0.15)])
    Rt2 = np.c_[R2 @ Rt1[:, :3], R2 @ Rt1[:, 3] + t2]
    Xw = np.c_[r.uniform(-4, 4, n), r.uniform(-3, 3, n), r.uniform(4, 14, n)]
    Xh = np.c_[Xw, np.ones(n)]

    def project(K, Rt):
        p = (K @ Rt @ Xh.T).T
        return p[:, :2] / p[:, 2:]
    p1 = project(K1, Rt1) + r.normal(0, noise, (n, 2))
    p2 = project(K2, Rt2) + r.normal(0, noise, (n, 2))
    bad = r.random(n) < outliers
    p2[bad] = np.c_[r.uniform(0, w, bad.sum()), r.uniform(0, h, bad.sum())]
    return p1.astype(np.float32).astype(np.float64), p2.astype(np.float32).astype(np.float64), Rt1, Rt2, K1, K2, R2, t2


def triangulate_cv2(cv2, p1, p2, Rt1, Rt2, K1, K2):
    """src/CameraPoseEstimator.cpp:86-132, one cv2 call per reference call."""
    P1, P2 = K1 @ Rt1, K2 @ Rt2
    X = np.zeros((len(p1), 3))
    front = np.zeros(len(p1), np.uint8)
    for i in range(len(p1)):
        A = np.array([P1[0] - P1[2] * p1[i, 0], P1[1] - P1[2] * p1[i, 1], P2[0] - P2[2] * p2[i, 0], P2[1] - P2[2] * p2[i, 1]])
        _, _, vt = cv2.SVDecomp(A, flags=cv2.SVD_MODIFY_A)
        x = vt[3] / vt[3, 3]
        X[i] = x[:3]
        front[i] = (Rt1 @ x)[2] > 0 and (Rt2 @ x)[2] > 0
    return X, front


# name -> (seed, n, noise px, outlier fraction, general first camera)
CASES = {
    "n50_clean": (1, 50, 0.0, 0.0, False),
    "n200": (2, 200, 0.5, 0.0, False),
    "n500_outliers": (3, 500, 0.5, 0.3, False),
    "n1000": (4, 1000, 0.7, 0.1, False),
    "n2000_general": (5, 2000, 0.5, 0.2, True),
    "n300_general_noisy": (6, 300, 1.5, 0.4, True),
    "n1": (7, 1, 0.3, 0.0, False),
}


def main():
    import cv2
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, (seed, n, noise, outl, general) in CASES.items():
        p1, p2, Rt1, Rt2, K1, K2, R2, t2 = scene(seed, n, noise, outl, general_first=general)
        X, front = triangulate_cv2(cv2, p1, p2, Rt1, Rt2, K1, K2)
        Xtp = cv2.triangulatePoints(K1 @ Rt1, K2 @ Rt2, p1.T.copy(), p2.T.copy())
        Xtp = (Xtp[:3] / Xtp[3]).T
        for k, v in (("p1", p1), ("p2", p2), ("Rt1", Rt1), ("Rt2", Rt2), ("K1", K1), ("K2", K2), ("X", X), ("front", front), ("Xtp", Xtp)):
            out["%s/%s" % (name, k)] = v
        # the bootstrap's hypothesis test (:326-349): R, R' and +-t candidates against the identity camera
        if not general:
            Rb = rot(0.3, 3.0, -0.2) @ R2          # a second, wrong rotation (stands in for the twisted-pair solution)
            Rts = np.stack([np.c_[R2, t2], np.c_[R2, -t2], np.c_[Rb, t2], np.c_[Rb, -t2]])
            order = np.random.default_rng(seed).permutation(4)
            Rts = Rts[order]
            counts = np.array([triangulate_cv2(cv2, p1, p2, Rt1, Rts[h], K1, K2)[1].sum() for h in range(4)], np.int32)
            out["%s/hyp_Rts" % name] = Rts
            out["%s/hyp_counts" % name] = counts
            out["%s/hyp_best" % name] = np.array(int(np.argmax(counts)), np.int32)     # first maximum == strict '<' update (:344)
        print(name, "n", n, "front", int(front.sum()), "max |X - Xtp| / |X| (inliers in front)",
              float(np.max(np.linalg.norm(X - Xtp, axis=1)[front > 0] / np.linalg.norm(X, axis=1)[front > 0])) if front.any() else None)
    np.savez_compressed(os.path.join(HERE, "tri_cases.npz"), **out)


if __name__ == "__main__":
    main()
