#!/usr/bin/env python
"""Golden vectors for 3-channel input (tests/golden/bgr_frame.npz), made like make_golden.py: cv2 4.13.0 in the build
container.  The reference hands OpenCV whatever imread(..., CV_LOAD_IMAGE_UNCHANGED) returned (src/FrameLoader.cpp:62), so
cv::ORB converts BGR frames to gray itself; this fixture pins that conversion and the extraction that follows it.

Run (build container only; needs cv2):  python tests/golden/make_golden_bgr.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from make_golden import cv_extract, sha  # noqa: E402
from monocular_slam_b200 import synthetic as syn  # noqa: E402


def main():
    import cv2
    cv2.setNumThreads(1)
    img = syn.bgr_frame(11, 333, 257)          # odd width: rows of 999 bytes, nothing aligned
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    k, d = cv_extract(cv2, img, 500, 0)        # ORB on the BGR frame itself
    kg, dg = cv_extract(cv2, gray, 500, 0)
    assert np.array_equal(k, kg) and np.array_equal(d, dg)
    big = syn.bgr_frame(12, 640, 480)
    kb, db = cv_extract(cv2, big, 1000, 0)
    # every (b, g, r) on a coarse lattice plus the extremes: pins the rounding of the fixed-point conversion
    v = np.array([0, 1, 2, 3, 17, 64, 127, 128, 129, 200, 253, 254, 255], np.uint8)
    lattice = np.stack(np.meshgrid(v, v, v, indexing="ij"), axis=-1).reshape(1, -1, 3)
    np.savez_compressed(os.path.join(HERE, "bgr_frame.npz"), img=img, gray_sha=np.array(sha(gray)), kp=k, desc=d,
                        big_sha=np.array(sha(big)), big_gray_sha=np.array(sha(cv2.cvtColor(big, cv2.COLOR_BGR2GRAY))), big_kp=kb, big_desc=db,
                        lattice=lattice, lattice_gray=cv2.cvtColor(lattice, cv2.COLOR_BGR2GRAY), cv2_version=np.array(cv2.__version__))
    print("wrote bgr_frame.npz: %d + %d keypoints" % (len(k), len(kb)))


if __name__ == "__main__":
    main()
