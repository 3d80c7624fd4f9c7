#!/usr/bin/env python
"""Device time of the fundamental-matrix outlier filter (k_fm_ransac) on BASELINE-shaped batches (the CPU baselines beside it
are in bench.py's `fundamental` section).  Usage: python tools/fmat_probe.py [npairs] [matches per pair] [inlier ratio]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monocular_slam_b200 import FundamentalFilter  # noqa: E402
from monocular_slam_b200 import synthetic as syn  # noqa: E402


def main():
    npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 320
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    inl = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
    cap = n
    p1 = np.zeros((npairs, cap, 2), np.float32)
    p2 = np.zeros((npairs, cap, 2), np.float32)
    for i in range(npairs):
        p1[i], p2[i] = syn.two_view_matches(100 + i, n, inl, 0.5)
    dev = torch.device("cuda:0")
    d1, d2 = torch.from_numpy(p1).to(dev), torch.from_numpy(p2).to(dev)
    dc = torch.full((npairs,), n, dtype=torch.int32, device=dev)
    ds = torch.zeros((npairs, cap), dtype=torch.uint8, device=dev)
    dF = torch.zeros((npairs, 9), dtype=torch.float64, device=dev)
    di = torch.zeros((npairs, 4), dtype=torch.int32, device=dev)
    fm = FundamentalFilter()
    stream = torch.cuda.Stream()
    fm.set_stream(stream.cuda_stream)
    for conf in (0.85, 0.99):
        with torch.cuda.stream(stream):
            for _ in range(3):
                fm.find_batch_dev(d1.data_ptr(), d2.data_ptr(), dc.data_ptr(), npairs, cap, 3.0, conf, ds.data_ptr(), dF.data_ptr(), di.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record(stream)
            for _ in range(reps):
                fm.find_batch_dev(d1.data_ptr(), d2.data_ptr(), dc.data_ptr(), npairs, cap, 3.0, conf, ds.data_ptr(), dF.data_ptr(), di.data_ptr())
            e1.record(stream)
            stream.synchronize()
        ms = e0.elapsed_time(e1) / reps
        info = di.cpu().numpy()
        print("conf %.2f: %d pairs x %d matches (inlier ratio %.2f): %.3f ms per batch = %.0f pairs/s; iterations mean %.1f max %d, "
              "candidates scored mean %.1f" % (conf, npairs, n, inl, ms, npairs / ms * 1e3, info[:, 1].mean(), info[:, 1].max(), info[:, 2].mean()))


if __name__ == "__main__":
    main()
