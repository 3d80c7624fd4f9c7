"""Small pass over every kernel for compute-sanitizer (gpurun only): extraction at two sizes and both score types, BGR,
compute() on border keypoints, batched + pipelined sequence matching, plain / split-K / chunk-merged / peer-memory kNN2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monocular_slam_b200 import (ORB, BFMatcher, FAST_SCORE, HARRIS_SCORE, KEYPOINT_DTYPE, DMATCH_DTYPE, CAMERAS_DTYPE, FundamentalFilter,
                                 Triangulator, _lib)
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
for mode in (_lib.KERNEL_INTEGER, _lib.KERNEL_TENSOR):          # XOR + POPC kernel, tcgen05 kernel
    m.set_kernel(mode)
    for nq, nt in [(1, 1), (300, 1000), (2000, 2000), (5, 70000), (1000, 3), (257, 129)]:
        q = syn.descriptors(1, nq); t = syn.descriptors(2, nt)
        m.knnMatch(q, t, 2); m.match_ratio(q, t, 0.8)
    # loop-closure scoring and NBestMatches
    fr = np.stack([syn.descriptors(10 + i, 300) for i in range(5)])
    m.loop_score(syn.descriptors(9, 200), fr, [300, 150, 0, 299, 1], 10, 100)
    m.NBestMatches(syn.descriptors(9, 70), syn.descriptors(8, 90), 10)
# batched frame pairs on the tensor-core kernel (caps large enough for the size-based choice) + outlier filter + triangulation
m.set_kernel(_lib.KERNEL_TENSOR)
orb = ORB(nfeatures=300, max_size=(640, 480), max_batch=4)
fm, tri = FundamentalFilter(), Triangulator()
kps, desc, cnt = orb.extract_batch(list(seq))
orb.match_consecutive(m, 0.8); orb.filter_consecutive(fm)
kps, desc, cnt = orb.extract_batch(list(seq))
good, ngood = orb.match_back(m, 3, 0.8); orb.filter_back(fm, 3)
p1, p2 = syn.two_view_matches(5, 300, 0.7, 0.5)
st, F, ni = fm.find_batch(p1[None], p2[None], [300])
cam = np.zeros(1, CAMERAS_DTYPE)
cam["Rt1"][0], cam["Rt2"][0] = np.c_[np.eye(3), np.zeros(3)], np.c_[np.eye(3), np.array([0.4, 0.05, 0.1])]
cam["K1"][0] = cam["K2"][0] = np.array([[900.0, 0, 960], [0, 900.0, 540], [0, 0, 1]])
tri.triangulate_batch(p1[None], p2[None], [300], cam, select=st)
tri.triangulate(p1.astype(np.float64), p2.astype(np.float64), cam["Rt1"][0], cam["Rt2"][0], cam["K1"][0], cam["K2"][0])
tri.triangulate_hypotheses(p1.astype(np.float64), p2.astype(np.float64), cam["Rt1"][0], np.stack([cam["Rt2"][0]] * 4), cam["K1"][0], cam["K2"][0])
# association / new-point selection on device-resident lists
back, cap, nprob = 3, 64, 2
g = np.zeros((nprob, back, cap), DMATCH_DTYPE)
ng = np.zeros((nprob, back), np.int64)
for p_ in range(nprob):
    for l in range(back):
        k_ = 40 + l
        g[p_, l, :k_]["query_idx"] = np.arange(k_); g[p_, l, :k_]["train_idx"] = (np.arange(k_) * 7) % cap
        ng[p_, l] = k_
d_g = torch.from_numpy(g.view(np.int32).reshape(nprob, back, cap, 4)).cuda(); d_ng = torch.from_numpy(ng).cuda()
d_pm = torch.from_numpy(np.where(np.arange(nprob * back * cap).reshape(nprob, back, cap) % 3 == 0, 5, -1).astype(np.int32)).cuda()
d_nc = torch.full((nprob,), 50, dtype=torch.int32, device="cuda")
d_cur = torch.zeros((nprob, cap), dtype=torch.int32, device="cuda"); d_aq = torch.zeros_like(d_cur); d_am = torch.zeros_like(d_cur)
d_na = torch.zeros(nprob, dtype=torch.int32, device="cuda"); d_acc = torch.zeros((nprob, back, cap), dtype=torch.uint8, device="cuda")
tri.associate_dev(d_g.data_ptr(), d_ng.data_ptr(), 0, d_pm.data_ptr(), d_nc.data_ptr(), nprob, back, cap, d_cur.data_ptr(), d_aq.data_ptr(), d_am.data_ptr(), d_na.data_ptr())
tri.select_new_dev(d_g.data_ptr(), d_ng.data_ptr(), 0, d_pm.data_ptr(), d_cur.data_ptr(), d_nc.data_ptr(), 0, nprob, back, cap, d_acc.data_ptr(), d_na.data_ptr())
tri.synchronize()
tri.close(); fm.close(); orb.close()
m.set_kernel(_lib.KERNEL_AUTO)
# JPEG ingest: every committed file (restart intervals 0 / 1 / 8 / 42, two table kinds), batched per image size
from monocular_slam_b200 import JpegDecoder
G = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "jpeg_cases.npz"))
jd = JpegDecoder()
by_shape = {}
for k_ in G["names"]:
    by_shape.setdefault(G[k_ + "_pixels"].shape, []).append(G[k_ + "_file"].tobytes())
for fl in by_shape.values():
    jd.decode(fl)
jd.close()
# bag of words: 16- and 32-lane descents, both key widths of the sort, every scoring type
from monocular_slam_b200 import Vocabulary
for shape in (dict(k=10, L=3), dict(k=19, L=2, ragged=True), dict(k=40, L=2)):
    va = syn.vocabulary(3, **shape)
    for scoring in range(6):
        voc = Vocabulary(va, scoring, scoring % 4)
        d = np.stack([syn.vocabulary_features(4, va, 300, pool=60), syn.descriptors(6, 300)])
        bows = voc.transform_batch(d, [300, 77], 1)
        voc.transform_batch(d, [0, 300])
        voc.score_batch(bows[0][0], [b[0] for b in bows] * 5)
        voc.transform_features(d[0], 2)
        voc.close()
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
