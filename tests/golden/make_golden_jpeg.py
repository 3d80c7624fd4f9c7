#!/usr/bin/env python
"""Golden vectors for the grey-scale JPEG decoder (tests/golden/jpeg_cases.npz).

The reference loads frames with cv::imread (src/FrameLoader.cpp:62); for JPEG files that is libjpeg behind OpenCV -- not part
of /root/reference.  The runnable instance here is cv2 4.13.0 (libjpeg-turbo 3.1.2, JDCT_ISLOW).  This script encodes seeded
grey frames with cv2.imencode under the settings that change the bit stream -- quality (the quantisation table), restart
interval, optimised Huffman tables, sizes that are not multiples of 8, a flat and a noise image -- and stores each file's
bytes together with what cv2.imdecode(file, IMREAD_UNCHANGED) returns; the same for colour files (4:2:0, 4:2:2, 4:4:4).
A progressive and a 4:1:1 file are stored too: the decoder must refuse them.

Run (build container only; needs cv2):  python tests/golden/make_golden_jpeg.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from monocular_slam_b200 import synthetic as syn  # noqa: E402


def main():
    out = {}
    rng = np.random.default_rng(3)
    images = {
        "tex333": syn.frame(3, 333, 257),                    # neither dimension a multiple of 8
        "nat200": syn.natural_frame(4, 200, 152),
        "noise64": rng.integers(0, 256, (48, 64), dtype=np.uint8),
        "flat17": np.full((9, 17), 201, np.uint8),
        "tiny1": np.array([[37]], np.uint8),
        "sat40": np.tile(np.array([[0, 255], [255, 0]], np.uint8), (20, 20)),   # full-swing checkerboard: values that hit the range limiter
    }
    settings = {
        "q90": [cv2.IMWRITE_JPEG_QUALITY, 90],
        "q50rst8": [cv2.IMWRITE_JPEG_QUALITY, 50, cv2.IMWRITE_JPEG_RST_INTERVAL, 8],
        "q100rst1": [cv2.IMWRITE_JPEG_QUALITY, 100, cv2.IMWRITE_JPEG_RST_INTERVAL, 1],
        "q75opt": [cv2.IMWRITE_JPEG_QUALITY, 75, cv2.IMWRITE_JPEG_OPTIMIZE, 1],
        "q10rst42opt": [cv2.IMWRITE_JPEG_QUALITY, 10, cv2.IMWRITE_JPEG_RST_INTERVAL, 42, cv2.IMWRITE_JPEG_OPTIMIZE, 1],
    }
    names = []
    for iname, img in images.items():
        for sname, params in settings.items():
            ok, enc = cv2.imencode(".jpg", img, params)
            assert ok
            dec = cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)
            assert dec.shape == img.shape and dec.dtype == np.uint8
            key = "%s_%s" % (iname, sname)
            out[key + "_file"] = enc.reshape(-1).copy()
            out[key + "_pixels"] = dec
            names.append(key)
    out["names"] = np.array(names)
    # colour files (YCbCr, interleaved scan): the three chroma samplings libjpeg writes by default or on request, odd sizes,
    # restart intervals counted in MCUs
    cnames = []
    colour = {"bgr71": syn.bgr_frame(5, 71, 53), "bgr16": syn.bgr_frame(6, 16, 16), "bgr1": syn.bgr_frame(7, 16, 16)[:1, :1].copy(),
              "noise33": rng.integers(0, 256, (33, 47, 3), dtype=np.uint8),
              "narrow3": rng.integers(0, 256, (21, 3, 3), dtype=np.uint8)}     # chroma two samples wide: libjpeg replicates instead of filtering
    csettings = {"420q90": [cv2.IMWRITE_JPEG_QUALITY, 90], "420q30rst3": [cv2.IMWRITE_JPEG_QUALITY, 30, cv2.IMWRITE_JPEG_RST_INTERVAL, 3],
                 "422q85rst1": [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_RST_INTERVAL, 1],
                 "444q100opt": [cv2.IMWRITE_JPEG_QUALITY, 100, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_OPTIMIZE, 1]}
    for iname, img in colour.items():
        for sname, params in csettings.items():
            ok, enc = cv2.imencode(".jpg", img, params)
            assert ok
            dec = cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)
            assert dec.shape == img.shape
            key = "%s_%s" % (iname, sname)
            out[key + "_file"] = enc.reshape(-1).copy()
            out[key + "_pixels"] = dec
            cnames.append(key)
    out["colour_names"] = np.array(cnames)
    ok, enc = cv2.imencode(".jpg", colour["bgr71"], [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411])
    out["refuse_411_file"] = enc.reshape(-1).copy()
    ok, enc = cv2.imencode(".jpg", images["tex333"], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    out["refuse_progressive_file"] = enc.reshape(-1).copy()
    path = os.path.join(HERE, "jpeg_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d files, %.0f KB" % (path, len(names), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
