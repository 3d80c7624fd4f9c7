#!/usr/bin/env python
"""Random soak of the GPU JPEG decoder against cv2.imdecode (run by hand on a GPU box that has cv2; not collected by pytest):
random sizes, qualities 1-100, restart intervals, optimised tables, image kinds, grey and colour (4:2:0 / 4:2:2 / 4:4:4); single
files and batches.
Usage: python tests/soak_jpeg.py [cases] [seed]"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from monocular_slam_b200 import JpegDecoder  # noqa: E402
from monocular_slam_b200 import synthetic as syn  # noqa: E402


def image(rng, w, h):
    kind = rng.integers(0, 6)
    if kind == 0:
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == 1:
        return np.full((h, w), rng.integers(0, 256), np.uint8)
    if kind == 2:
        return (np.add.outer(np.arange(h), np.arange(w)) * rng.integers(1, 5) % 256).astype(np.uint8)
    if kind == 3 and w >= 64 and h >= 64:
        return syn.natural_frame(int(rng.integers(1 << 30)), w, h)
    if kind == 4 and w >= 64 and h >= 64:
        return syn.frame(int(rng.integers(1 << 30)), w, h)
    img = np.zeros((h, w), np.uint8)                 # sparse bright dots on black: long runs of zero blocks (a periodic stream)
    img[rng.integers(0, h, 5), rng.integers(0, w, 5)] = 255
    return img


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    dec = JpegDecoder()
    bad = 0
    for c in range(cases):
        big = rng.random() < 0.15
        w, h = (int(rng.integers(300, 2000)), int(rng.integers(300, 1200))) if big else (int(rng.integers(1, 300)), int(rng.integers(1, 300)))
        n = int(rng.integers(1, 5))
        files, refs = [], []
        colour = rng.random() < 0.4
        samp = int(rng.choice([cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]))
        for _ in range(n):
            q = int(rng.integers(1, 101))
            rst = int(rng.choice([0, 0, 1, 2, 5, max(1, (w + 7) // 8), 100, 1000]))
            params = [cv2.IMWRITE_JPEG_QUALITY, q] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else []) + \
                     ([cv2.IMWRITE_JPEG_OPTIMIZE, 1] if rng.random() < 0.3 else [])
            img = image(rng, w, h)
            if colour:                                   # three different planes; one sampling per batch
                img = np.stack([img, image(rng, w, h), np.roll(img, 3, axis=1)], -1)
                params = params + [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, samp]
            f = cv2.imencode(".jpg", img, params)[1].tobytes()
            files.append(f)
            refs.append(cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_UNCHANGED))
        got = dec.decode(files)
        for i in range(n):
            if not np.array_equal(got[i], refs[i]) or (c % 10 == 0 and not np.array_equal((oracle.jpeg_decode_bgr if colour else oracle.jpeg_decode_gray)(files[i]), refs[i])):
                bad += 1
                print("case %d file %d (%dx%d): MISMATCH (%d pixels)" % (c, i, w, h, int((got[i] != refs[i]).sum())))
    dec.close()
    print("soak_jpeg: %d cases, %d mismatches" % (cases, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
