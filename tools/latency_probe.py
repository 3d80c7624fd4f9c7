"""Scratch: single-frame latency of the drop-in call pattern (detect, then compute) through host buffers (gpurun only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monocular_slam_b200 import ORB, BFMatcher
from monocular_slam_b200 import synthetic as syn
seq = syn.sequence(4, 1920, 1080, seed=1)
orb = ORB(nfeatures=2000, max_size=(1920, 1080), max_batch=1)
m = BFMatcher()
def t(fn, n=50):
    for _ in range(5): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
img = seq[0]
print("detect            %.3f ms" % t(lambda: orb.detect(img)))
k = orb.detect(img)
print("compute           %.3f ms" % t(lambda: orb.compute(img, k)))
print("detectAndCompute  %.3f ms" % t(lambda: orb.detectAndCompute(img)))
k0, d0 = orb.detectAndCompute(seq[0]); k1, d1 = orb.detectAndCompute(seq[1])
print("match_ratio 2kx2k %.3f ms" % t(lambda: m.match_ratio(d1, d0, 0.8)))
print("knnMatch 2kx2k    %.3f ms" % t(lambda: m.knnMatch(d1, d0, 2)))
pin = torch.from_numpy(img).pin_memory().numpy()
print("detectAndCompute (pinned frame) %.3f ms" % t(lambda: orb.detectAndCompute(pin)))
# the call after matching: computeFundamentalMatrix of one pair (1000 matches, 70 % inliers), the reference's argument list
from monocular_slam_b200 import DMATCH_DTYPE, KEYPOINT_DTYPE, FundamentalFilter
fm = FundamentalFilter()
p1, p2 = syn.two_view_matches(3, 1000, 0.7, 0.5)
ka = np.zeros(1000, KEYPOINT_DTYPE); kb = np.zeros(1000, KEYPOINT_DTYPE)
ka["x"], ka["y"], kb["x"], kb["y"] = p1[:, 0], p1[:, 1], p2[:, 0], p2[:, 1]
mt = np.zeros(1000, DMATCH_DTYPE); mt["query_idx"] = mt["train_idx"] = np.arange(1000)
print("computeFundamentalMatrix 1 pair x 1000 matches %.3f ms" % t(lambda: fm.compute_fundamental(ka, kb, mt)))
