#!/usr/bin/env python
"""Soak comparison of the ORB extraction and the Hamming matcher on the GPU against the CPU oracle: random frame sizes,
textures and runtime parameters (keypoints, responses, angles, descriptors bit for bit), then random descriptor sets with
planted duplicates (kNN2 indices / distances and the ratio-filtered list).
Not collected by pytest (minutes of CPU oracle time): run by hand on a GPU box.
Usage: python tests/soak_orb.py [nframes] [nmatch_cases]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (the checker)
from monocular_slam_b200 import ORB, BFMatcher, OrbxError  # noqa: E402
from monocular_slam_b200 import synthetic as syn  # noqa: E402


def main():
    nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    nmatch = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    r = np.random.default_rng(777)
    bad = skipped = 0
    t0 = time.time()
    for i in range(nframes):
        w, h = int(r.integers(70, 1500)), int(r.integers(70, 900))
        nf = int(r.choice([100, 500, 1000, 2000, 4000]))
        sf = float(r.choice([1.05, 1.1, 1.2, 1.2, 1.5, 2.0]))
        nl = int(r.integers(1, 13))
        st = int(r.integers(0, 2))
        thr = int(r.choice([5, 10, 20, 20, 40, 80]))
        seed = int(r.integers(1 << 30))
        kind = int(r.integers(0, 5))
        if kind == 0:
            img = syn.frame(seed, w, h, nrect=max(5, w * h // 4000))
        elif kind == 1:
            img = syn.textured_frame(seed, w, h)
        elif kind == 2:
            img = syn.natural_frame(seed, w, h)          # camera-like: K2's compass pre-test and work lists do the scoring
        elif kind == 3:                                  # half camera-like, half corner-dense: per-row decisions differ inside a tile
            img = syn.natural_frame(seed, w, h)
            img[:, w // 2:] = syn.frame(seed + 1, w, h, nrect=max(5, w * h // 4000))[:, w // 2:]
        else:
            img = np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)      # white noise: corners everywhere
        P = oracle.Params(nfeatures=nf, scale_factor=sf, nlevels=nl, score_type=st, fast_threshold=thr)
        try:
            orb = ORB(nfeatures=nf, scaleFactor=sf, nlevels=nl, scoreType=st, fastThreshold=thr, max_size=(w, h), max_batch=1)
        except OrbxError as e:           # a pyramid level of this size / scale would be empty: rejected like cv2 rejects a 0-size resize
            skipped += 1
            assert "is empty" in str(e), e
            continue
        ok, od = oracle.detect_and_compute(img, P)
        k, d = orb.detectAndCompute(img)
        if r.random() < 0.5:             # again on the same handle: K3's density hint of the first call now steers K2
            k, d = orb.detectAndCompute(img)
        orb.close()
        same = len(k) == len(ok) and all(np.array_equal(k[f], ok[f]) for f in ok.dtype.names) and np.array_equal(d, od)
        if not same:
            bad += 1
            print("EXTRACT MISMATCH: %dx%d nf %d sf %.2f nl %d score %d thr %d seed %d kind %d: %d vs %d keypoints"
                  % (w, h, nf, sf, nl, st, thr, seed, kind, len(k), len(ok)))
    print("%d frames compared in %.0f s (%d configurations rejected for an empty pyramid level): %d mismatches" % (nframes - skipped, time.time() - t0, skipped, bad))
    from monocular_slam_b200 import _lib
    m = BFMatcher()
    badm = 0
    t0 = time.time()
    for i in range(nmatch):
        nq, nt = int(r.integers(0, 3000)), int(r.integers(0, 3000))
        q = r.integers(0, 256, (nq, 32), dtype=np.uint8)
        t = r.integers(0, 256, (nt, 32), dtype=np.uint8)
        if nt > 4 and nq > 4:                        # duplicates and near-duplicates: ties must go to the lowest train index
            t[r.integers(0, nt, nt // 5)] = t[r.integers(0, nt, nt // 5)]
            src = r.integers(0, nt, nq // 2)
            q[: nq // 2] = t[src]
            flips = r.integers(0, 256, (nq // 2, 32), dtype=np.uint8) & r.integers(0, 256, (nq // 2, 32), dtype=np.uint8) & r.integers(0, 256, (nq // 2, 32), dtype=np.uint8)
            q[: nq // 2] ^= flips * (r.random((nq // 2, 1)) < 0.7).astype(np.uint8)
        ratio = float(r.choice([0.6, 0.75, 0.8, 0.85, 1.0]))
        m.set_kernel(int(r.choice([_lib.KERNEL_AUTO, _lib.KERNEL_INTEGER, _lib.KERNEL_TENSOR])))    # XOR + POPC or tcgen05 contraction
        good = m.match_ratio(q, t, ratio)
        gq, gt, gd = oracle.match_features(q, t, ratio)
        if not (np.array_equal(good["query_idx"], gq) and np.array_equal(good["train_idx"], gt) and np.array_equal(good["distance"].astype(np.int32), gd)):
            badm += 1
            print("MATCH MISMATCH: nq %d nt %d ratio %.2f" % (nq, nt, ratio))
    print("%d matcher cases compared in %.0f s: %d mismatches" % (nmatch, time.time() - t0, badm))
    m.close()


if __name__ == "__main__":
    main()
