#!/usr/bin/env python
"""Golden vectors for the whole chain on image data (tests/golden/chain_cases.npz): the reference's bootstrap sequence
(src/CameraPoseEstimator.cpp:270-291) on two views of a layered-depth scene (monocular_slam_b200/synthetic.py layered_pair):
    detect + compute on both frames                         (FeatureExtractor.cpp:17,19)
    matchFeatures(desc_cur, desc_prev, ratio)               (CameraPoseEstimator.cpp:200-213)
    findFundamentalMat(FM_RANSAC, 3, 0.85, status)          (:563)
    findFundamentalMat(inliers, FM_8POINT)                  (:585)
all through cv2 4.13.0.  Keypoints are put in the canonical (octave, y, x) order of include/orbx.h before matching, so that
the match list -- and with it RANSAC's sample sequence -- is the one the replacement produces.

Run (build container only; needs cv2):  python tests/golden/make_golden_chain.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from make_golden import cv_extract, cv_knn2, cv_ratio  # noqa: E402
from monocular_slam_b200 import synthetic as syn  # noqa: E402

# name -> (seed, w, h, layers, motion (dy, dx), nfeatures, ratio)
CASES = {
    "kitti_like": (1, 1241, 376, 4, (2, 5), 2000, 0.8),
    "vga_5layers": (2, 800, 600, 5, (3, 4), 1500, 0.8),
    "hd_ratio75": (3, 1920, 1080, 6, (2, 3), 2000, 0.75),
}


def main():
    import cv2
    cv2.setNumThreads(1)
    out = {"cv2_version": np.array(cv2.__version__), "names": np.array(list(CASES))}
    for name, (seed, w, h, layers, motion, nf, ratio) in CASES.items():
        prev, cur = syn.layered_pair(seed, w, h, layers, motion)
        kp_prev, d_prev = cv_extract(cv2, prev, nf, 0)
        kp_cur, d_cur = cv_extract(cv2, cur, nf, 0)
        idx, dist = cv_knn2(cv2, d_cur, d_prev)
        good = cv_ratio(idx, dist, ratio)                      # rows (queryIdx, trainIdx, distance), ascending queryIdx
        p1 = np.stack([kp_cur["x"][good[:, 0]], kp_cur["y"][good[:, 0]]], 1).astype(np.float64)
        p2 = np.stack([kp_prev["x"][good[:, 1]], kp_prev["y"][good[:, 1]]], 1).astype(np.float64)
        F, mask = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 3.0, 0.85)
        mask = mask.ravel().astype(np.uint8)
        F8, _ = cv2.findFundamentalMat(p1[mask > 0], p2[mask > 0], cv2.FM_8POINT)
        assert F8 is not None and F8.shape == (3, 3)
        out[name + "_cfg"] = np.array([seed, w, h, layers, motion[0], motion[1], nf, ratio], np.float64)
        out[name + "_counts"] = np.array([len(kp_prev), len(kp_cur), len(good), int(mask.sum())], np.int32)
        out[name + "_good"] = good
        out[name + "_mask"] = np.packbits(mask)
        out[name + "_F8"] = F8
        print("%-12s %4d / %4d keypoints, %4d matches, %4d inliers" % (name, len(kp_prev), len(kp_cur), len(good), int(mask.sum())))
    path = os.path.join(HERE, "chain_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
