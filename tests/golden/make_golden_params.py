#!/usr/bin/env python
"""Golden vectors for non-default runtime parameters (tests/golden/params_cases.npz): nlevels, scaleFactor, fastThreshold,
nfeatures and scoreType are arguments of the replacement (include/orbx.h orbx_params); the reference default-constructs
cv::ORB (src/FeatureExtractor.h:23-24) but BASELINE.json asks for other feature counts, so the whole parameter surface
that orbx_create accepts is pinned against cv2 4.13.0 here.

Run (build container only; needs cv2):  python tests/golden/make_golden_params.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from make_golden import KP  # noqa: E402
from monocular_slam_b200 import synthetic as syn  # noqa: E402

# name -> (nfeatures, scaleFactor, nlevels, scoreType, fastThreshold)
CASES = {
    "l4_s15": (800, 1.5, 4, 0, 20),
    "l1": (300, 1.2, 1, 0, 20),
    "l12_s11_t30": (1500, 1.1, 12, 0, 30),
    "t10_fast": (800, 1.2, 8, 1, 10),
    "l3_s20_t5": (600, 2.0, 3, 0, 5),
    "l16_s105": (2000, 1.05, 16, 1, 20),
}


def main():
    import cv2
    cv2.setNumThreads(1)
    img = syn.frame(9, 640, 480)
    out = {"img_sha": np.array(__import__("hashlib").sha256(img.tobytes()).hexdigest()), "cv2_version": np.array(cv2.__version__)}
    for name, (nf, sf, nl, st, thr) in CASES.items():
        orb = cv2.ORB_create(nfeatures=nf, scaleFactor=sf, nlevels=nl, scoreType=st, fastThreshold=thr)
        kp = orb.detect(img, None)
        kp, des = orb.compute(img, kp)
        a = np.zeros(len(kp), KP)
        for i, k in enumerate(kp):
            a[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
        if des is None:
            des = np.zeros((0, 32), np.uint8)
        o = np.lexsort((a["x"], a["y"], a["octave"]))
        out[name + "_kp"], out[name + "_desc"] = a[o], des[o]
        out[name + "_params"] = np.array([nf, sf, nl, st, thr], np.float64)
        print(name, len(kp), "keypoints; per octave", np.bincount(a["octave"], minlength=nl).tolist())
    np.savez_compressed(os.path.join(HERE, "params_cases.npz"), **out)


if __name__ == "__main__":
    main()
