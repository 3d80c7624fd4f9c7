#!/usr/bin/env python
"""Golden vectors for the fundamental-matrix outlier filter (tests/golden/fmat_cases.npz).

computeFundamentalMatrix (reference src/CameraPoseEstimator.cpp:545-586) is two OpenCV calls,
findFundamentalMat(FM_RANSAC, 3, 0.85, status) and findFundamentalMat(inliers, FM_8POINT).  OpenCV is not vendored in the
reference; as for the ORB path the runnable cv2 4.13.0 of this container is the parity target, and its outputs on seeded
synthetic two-view match sets (monocular_slam_b200/synthetic.py two_view_matches) are committed here.

Run (build container only; needs cv2):  python tests/golden/make_golden_fmat.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from monocular_slam_b200 import synthetic as syn  # noqa: E402

# name -> (seed, n, inlier ratio, noise px, max_distance, confidence)
CASES = {
    "n15": (1, 15, 0.9, 0.3, 3.0, 0.85),
    "n40_clean": (2, 40, 0.85, 0.3, 3.0, 0.99),
    "n128": (3, 128, 0.7, 0.5, 3.0, 0.85),
    "n300_low": (4, 300, 0.35, 0.5, 3.0, 0.99),          # runs the full 1000 iterations
    "n500": (5, 500, 0.6, 0.5, 3.0, 0.85),
    "n800_tight": (6, 800, 0.75, 0.4, 1.0, 0.85),
    "n1000": (7, 1000, 0.7, 0.5, 3.0, 0.85),
    "n1200_99": (8, 1200, 0.5, 0.7, 3.0, 0.99),
    "n2000": (9, 2000, 0.8, 0.5, 3.0, 0.85),
    "n2000_loose": (10, 2000, 0.45, 1.0, 5.0, 0.85),
    "n64_defaults": (11, 64, 0.7, 0.5, 0.0, 0.0),         # OpenCV substitutes 3 / 0.99
    "n250_noise": (12, 250, 0.65, 1.5, 3.0, 0.85),
    # 8..14 points: OpenCV runs LMedS instead of RANSAC; reproducible only with 14 (see include/orbx.h)
    "n14_lmeds_a": (13, 14, 0.8, 0.4, 3.0, 0.85),
    "n14_lmeds_b": (14, 14, 0.9, 0.8, 3.0, 0.99),
    "n14_lmeds_c": (15, 14, 0.7, 0.3, 3.0, 0.85),
}


def main():
    import cv2
    cv2.setNumThreads(1)
    out = {"cv2_version": np.array(cv2.__version__), "names": np.array(list(CASES))}
    for name, (seed, n, inl, noise, thr, conf) in CASES.items():
        p1, p2 = syn.two_view_matches(seed, n, inl, noise)
        Fr, mask = cv2.findFundamentalMat(p1.astype(np.float64), p2.astype(np.float64), cv2.FM_RANSAC, thr, conf)
        assert Fr is not None and Fr.shape == (3, 3), name
        mask = mask.ravel().astype(np.uint8)
        a, b = p1[mask > 0].astype(np.float64), p2[mask > 0].astype(np.float64)
        F8 = cv2.findFundamentalMat(a, b, cv2.FM_8POINT)[0] if mask.sum() >= 8 else np.zeros((3, 3))
        assert F8 is not None and F8.shape == (3, 3), name
        out[name + "_cfg"] = np.array([seed, n, inl, noise, thr, conf], np.float64)
        out[name + "_mask"] = np.packbits(mask)
        out[name + "_Fransac"] = Fr
        out[name + "_F8"] = F8
        print("%-14s n=%4d inliers=%4d" % (name, n, int(mask.sum())))
    path = os.path.join(HERE, "fmat_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
