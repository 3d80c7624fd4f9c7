#!/usr/bin/env python
"""Soak comparison of the fundamental-matrix filter on the GPU against the CPU oracle: many random two-view match sets
(sizes, inlier ratios, noise, thresholds, confidences, image sizes), every status mask and iteration count compared.
Not collected by pytest (minutes of CPU oracle time): run by hand on a GPU box.
Usage: python tests/soak_fmat.py [npairs_total]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (the checker)
from monocular_slam_b200 import FundamentalFilter  # noqa: E402
from monocular_slam_b200 import synthetic as syn  # noqa: E402


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    fm = FundamentalFilter()
    r = np.random.default_rng(2026)
    done = bad = worst_f = 0
    t0 = time.time()
    while done < total:
        npairs, cap = 200, int(r.choice([64, 300, 700, 1500]))
        thr = float(r.choice([0.5, 1.0, 2.0, 3.0, 5.0]))
        conf = float(r.choice([0.85, 0.95, 0.99]))
        size = [(640, 480), (1241, 376), (1920, 1080), (3840, 2160)][int(r.integers(0, 4))]
        counts = r.integers(14, cap + 1, npairs).astype(np.int32)         # 14: the LMedS path
        p1 = np.zeros((npairs, cap, 2), np.float32)
        p2 = np.zeros((npairs, cap, 2), np.float32)
        for i in range(npairs):
            a, b = syn.two_view_matches(int(r.integers(1 << 30)), int(counts[i]), float(r.uniform(0.2, 0.99)), float(r.uniform(0.0, 2.5)), size)
            p1[i, :counts[i]], p2[i, :counts[i]] = a, b
        status, F, ninl = fm.find_batch(p1, p2, counts, thr, conf)
        info = fm.last_info(npairs)
        for i in range(npairs):
            n = int(counts[i])
            _, mo, iters = oracle.fm_ransac(p1[i, :n], p2[i, :n], thr, conf)
            if not np.array_equal(status[i, :n], mo) or info[i, 1] != iters:
                bad += 1
                os.makedirs("gpurun_out", exist_ok=True)
                np.savez("gpurun_out/soak_bad_%d.npz" % bad, p1=p1[i, :n], p2=p2[i, :n], thr=thr, conf=conf, gpu=status[i, :n], cpu=mo,
                         gpu_F=F[i], info=info[i])
                print("MISMATCH: n %d thr %.1f conf %.2f size %s: %d status bytes differ, iterations %d vs %d"
                      % (n, thr, conf, size, int((status[i, :n] != mo).sum()), info[i, 1], iters))
            elif mo.sum() >= 8:
                Fo = oracle.fm_8point(p1[i, :n][mo > 0], p2[i, :n][mo > 0])
                if Fo is not None:
                    worst_f = max(worst_f, float(np.linalg.norm(F[i] - Fo) / np.linalg.norm(Fo)))
        done += npairs
    print("%d pairs compared in %.0f s: %d mismatching status masks / iteration counts, worst relative F error %.2e"
          % (done, time.time() - t0, bad, worst_f))


if __name__ == "__main__":
    main()
