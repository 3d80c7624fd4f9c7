"""Small pass over every kernel for compute-sanitizer (gpurun only): extraction at two sizes and both score types, BGR,
compute() on border keypoints, batched + pipelined sequence matching, plain / split-K / chunk-merged / peer-memory kNN2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monocular_slam_b200 import ORB, BFMatcher, FAST_SCORE, HARRIS_SCORE, KEYPOINT_DTYPE, DMATCH_DTYPE
from monocular_slam_b200 import synthetic as syn

seq = syn.sequence(4, 333, 257, seed=5)
for st in (HARRIS_SCORE, FAST_SCORE):
    orb = ORB(nfeatures=300, scoreType=st, max_size=(640, 480), max_batch=4)
    k, d = orb.detectAndCompute(seq[0])
    k2 = orb.detect(seq[1]); k2, d2 = orb.compute(seq[1], k2)
    kps, desc, cnt = orb.extract_batch(list(seq))
    big = syn.frame(3, 640, 480)
    kk = orb.detect(big)
    kk["x"][:5] = 31.0; kk["y"][:5] = 31.0
    orb.compute(big, kk)
    orb.detectAndCompute(syn.bgr_frame(2, 333, 257))
    m = BFMatcher()
    good, ngood = orb.match_consecutive(m, 0.8, kps.shape[1], 4) if False else (None, None)
    kps, desc, cnt = orb.extract_batch(list(seq))
    good, ngood = orb.match_consecutive(m, 0.8, kps.shape[1], 4)
    cap = orb.default_cap
    outs = [(np.zeros((2, cap), KEYPOINT_DTYPE), np.zeros((2, cap, 32), np.uint8), np.zeros(2, np.int32), np.zeros((2, cap), DMATCH_DTYPE), np.zeros(2, np.int64)) for _ in range(3)]
    for i in range(3):
        orb.submit_batch(list(seq[:2]), m, 0.8, outs[i])
    while orb.batches_in_flight():
        orb.wait_batch()
    m.close(); orb.close()
m = BFMatcher()
for nq, nt in [(1, 1), (300, 1000), (2000, 2000), (5, 70000), (1000, 3)]:
    q = syn.descriptors(1, nq); t = syn.descriptors(2, nt)
    m.knnMatch(q, t, 2); m.match_ratio(q, t, 0.8)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ms = [BFMatcher() for _ in range(3)]
    for x in ms: x.set_stream(s.cuda_stream)
    bases = [x.p2p_export(400, 3, r)[1] for r, x in enumerate(ms)]
    for x in ms: x.p2p_import_ptrs(bases)
    dq = torch.from_numpy(syn.descriptors(3, 300)).cuda(); dt = torch.from_numpy(syn.descriptors(4, 900)).cuda()
    outs = [torch.empty((300, 4), dtype=torch.int32, device="cuda") for _ in range(3)]
    for rep in range(2):
        for r, x in enumerate(ms): x.knn2_p2p_scatter_dev(dq.data_ptr(), 300, dt[300 * r:300 * (r + 1)].contiguous().data_ptr(), 300, 300 * r)
        for r, x in enumerate(ms): x.p2p_merge_dev(300, outs[r].data_ptr())
    s.synchronize()
    for x in ms: x.close()
m.close()
print("sanitize smoke done")
