"""GPU: the fundamental-matrix outlier filter (monocular_slam_b200/csrc/fmat.cu through the C ABI) against the cv2 golden
vectors and the CPU oracle.  Status masks and inlier counts: identical.  F: relative Frobenius error <= 1e-6."""
import os

import numpy as np
import pytest

import oracle
from monocular_slam_b200 import DMATCH_DTYPE, KEYPOINT_DTYPE, BFMatcher, FundamentalFilter, ORB
from monocular_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "fmat_cases.npz"))
NAMES = [str(n) for n in GOLD["names"]]
F_RTOL = 1e-6


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def case(name):
    seed, n, inl, noise, thr, conf = GOLD[name + "_cfg"]
    p1, p2 = syn.two_view_matches(int(seed), int(n), float(inl), float(noise))
    return p1, p2, float(thr), float(conf), np.unpackbits(GOLD[name + "_mask"])[:int(n)], GOLD[name + "_F8"]


@pytest.fixture(scope="module")
def fm():
    f = FundamentalFilter()
    yield f
    f.close()


@pytest.mark.parametrize("name", NAMES)
def test_golden_single_pair(fm, name):
    p1, p2, thr, conf, mask, F8 = case(name)
    status, F, ninl = fm.find_batch(p1[None], p2[None], [len(p1)], thr, conf)
    assert np.array_equal(status[0], mask), "%s: %d status bytes differ" % (name, int((status[0] != mask).sum()))
    assert ninl[0] == mask.sum()
    assert rel(F[0], F8) <= F_RTOL
    info = fm.last_info(1)[0]
    assert info[0] == ninl[0] and 1 <= info[1] <= 1000 and info[3] == 1


def test_golden_cases_as_one_ragged_batch(fm):
    cases = [case(n) for n in NAMES if GOLD[n + "_cfg"][4] == 3.0 and GOLD[n + "_cfg"][5] == 0.85]
    cap = max(len(c[0]) for c in cases) + 37
    p1 = np.full((len(cases), cap, 2), np.nan, np.float32)        # slack past counts[] must never be read
    p2 = np.full((len(cases), cap, 2), np.nan, np.float32)
    for i, c in enumerate(cases):
        p1[i, :len(c[0])] = c[0]
        p2[i, :len(c[0])] = c[1]
    status, F, ninl = fm.find_batch(p1, p2, [len(c[0]) for c in cases], 3.0, 0.85)
    for i, c in enumerate(cases):
        n = len(c[0])
        assert np.array_equal(status[i, :n], c[4]) and not status[i, n:].any()
        assert rel(F[i], c[5]) <= F_RTOL


def test_reference_signature_with_keypoints_and_dmatches(fm):
    p1, p2, thr, conf, mask, F8 = case("n500")
    n = len(p1)
    perm = np.random.default_rng(2).permutation(n)
    k1 = np.zeros(n + 11, KEYPOINT_DTYPE)
    k2 = np.zeros(n + 5, KEYPOINT_DTYPE)
    k1["x"][perm], k1["y"][perm] = p1[:, 0], p1[:, 1]
    k2["x"][:n], k2["y"][:n] = p2[:, 0], p2[:, 1]
    m = np.zeros(n, DMATCH_DTYPE)
    m["query_idx"], m["train_idx"] = perm, np.arange(n)
    F, status, ninl = fm.compute_fundamental(k1, k2, m, thr, conf)
    assert np.array_equal(status, mask) and ninl == mask.sum() and rel(F, F8) <= F_RTOL
    Fo, so, no = oracle.compute_fundamental(np.c_[k1["x"], k1["y"]], np.c_[k2["x"], k2["y"]], m, thr, conf)
    assert np.array_equal(status, so) and ninl == no and rel(F, Fo) <= F_RTOL
    m["query_idx"][3] = n + 11
    with pytest.raises(Exception):
        fm.compute_fundamental(k1, k2, m, thr, conf)


def test_random_scenes_against_oracle(fm):
    r = np.random.default_rng(77)
    npairs, cap = 48, 1500
    counts = r.integers(15, cap + 1, npairs).astype(np.int32)
    p1 = np.zeros((npairs, cap, 2), np.float32)
    p2 = np.zeros((npairs, cap, 2), np.float32)
    for i in range(npairs):
        a, b = syn.two_view_matches(500 + i, int(counts[i]), float(r.uniform(0.3, 0.95)), float(r.uniform(0.2, 1.2)))
        p1[i, :counts[i]], p2[i, :counts[i]] = a, b
    for thr, conf in ((3.0, 0.85), (2.0, 0.99)):
        status, F, ninl = fm.find_batch(p1, p2, counts, thr, conf)
        info = fm.last_info(npairs)
        for i in range(npairs):
            n = int(counts[i])
            Fr, mo, iters = oracle.fm_ransac(p1[i, :n], p2[i, :n], thr, conf)
            assert np.array_equal(status[i, :n], mo), "pair %d: %d status bytes differ" % (i, int((status[i, :n] != mo).sum()))
            assert info[i, 1] == iters and ninl[i] == mo.sum()
            if mo.sum() >= 8:
                assert rel(F[i], oracle.fm_8point(p1[i, :n][mo > 0], p2[i, :n][mo > 0])) <= F_RTOL


def test_small_sets_degenerate_sets_and_empty_batch(fm):
    p1, p2 = syn.two_view_matches(1, 14, 1.0, 0.1)
    status, F, ninl = fm.find_batch(p1[None, :7], p2[None, :7], [7])   # < 8 correspondences: no model (see include/orbx.h)
    assert ninl[0] == 0 and not status.any() and not F.any()
    status, F, ninl = fm.find_batch(np.zeros((3, 64, 2), np.float32), np.zeros((3, 64, 2), np.float32), [0, 64, 20])
    assert not ninl.any() and not F.any()                               # identical points: every sample is rejected as collinear
    r = np.random.default_rng(5)
    a = r.uniform(0, 1000, (1, 300, 2)).astype(np.float32)
    b = r.uniform(0, 1000, (1, 300, 2)).astype(np.float32)
    status, F, ninl = fm.find_batch(a, b, [300])                        # pure outliers: 1000 iterations, same mask as the CPU
    Fo, mo, iters = oracle.fm_ransac(a[0], b[0], 3.0, 0.85)
    assert np.array_equal(status[0], mo) and fm.last_info(1)[0, 1] == iters == 1000
    status, F, ninl = fm.find_batch(np.zeros((0, 8, 2), np.float32), np.zeros((0, 8, 2), np.float32), np.zeros(0, np.int32))
    assert status.shape == (0, 8)
    with pytest.raises(Exception):
        fm.find_batch(a, b, [301])


def test_full_size_batch_properties(fm):
    """BASELINE-sized batch (64 frames x 5 predecessors, ~1000 matches per pair): permuting the pairs permutes the results,
    a second run is identical, and every reported inlier satisfies the epipolar test of the returned 8-point matrix's
    RANSAC parent (checked through the oracle on a sample of pairs)."""
    npairs, cap = 320, 1200
    r = np.random.default_rng(9)
    counts = r.integers(600, cap + 1, npairs).astype(np.int32)
    p1 = np.zeros((npairs, cap, 2), np.float32)
    p2 = np.zeros((npairs, cap, 2), np.float32)
    for i in range(npairs):
        a, b = syn.two_view_matches(2000 + i, int(counts[i]), 0.55 + 0.4 * (i % 7) / 7, 0.5)
        p1[i, :counts[i]], p2[i, :counts[i]] = a, b
    s1, F1, n1 = fm.find_batch(p1, p2, counts)
    s2, F2, n2 = fm.find_batch(p1, p2, counts)
    assert np.array_equal(s1, s2) and np.array_equal(F1, F2) and np.array_equal(n1, n2)
    perm = r.permutation(npairs)
    s3, F3, n3 = fm.find_batch(p1[perm], p2[perm], counts[perm])
    assert np.array_equal(s3, s1[perm]) and np.array_equal(F3, F1[perm]) and np.array_equal(n3, n1[perm])
    assert np.array_equal(s1.sum(1), n1) and (n1 >= 0.5 * 0.55 * counts).all()
    # a batch of more pairs than CTA slots leaves the eigen-decomposition of the 8-point step to a second kernel (k_fm_eight_point);
    # in two halves the same pairs run through the single kernel: bit-identical matrices, masks and counts
    half = npairs // 2
    for lo, hi in ((0, half), (half, npairs)):
        s4, F4, n4 = fm.find_batch(p1[lo:hi], p2[lo:hi], counts[lo:hi])
        assert np.array_equal(s4, s1[lo:hi]) and np.array_equal(F4, F1[lo:hi]) and np.array_equal(n4, n1[lo:hi])
    assert (np.abs(F1.reshape(npairs, -1)).sum(1) > 0).all()
    for i in range(0, npairs, 23):
        n = int(counts[i])
        Fo, mo, _ = oracle.fm_ransac(p1[i, :n], p2[i, :n], 3.0, 0.85)
        assert np.array_equal(s1[i, :n], mo)


def test_filter_consecutive_keeps_matches_on_device(fm):
    """orbx_filter_consecutive == fmx_compute_fundamental on the downloaded keypoints and match lists of every pair."""
    w, h, n, nf = 640, 480, 5, 800
    frames = syn.sequence(2 * n, w, h, seed=3)
    orb = ORB(nfeatures=nf, max_size=(w, h), max_batch=n)
    bf = BFMatcher()
    prev_kps = None
    for b in range(2):
        kps, desc, counts = orb.extract_batch(frames[b * n:(b + 1) * n])
        cap = kps.shape[1]
        good, ngood = orb.match_consecutive(bf, 0.8, cap, n)
        status, F, ninl = orb.filter_consecutive(fm, cap, n)
        for f in range(n):
            kt = kps[f - 1] if f else prev_kps
            if kt is None:
                assert ninl[f] == 0 and not status[f].any()
                continue
            Fh, sh, nh = fm.compute_fundamental(kps[f], kt, good[f, :ngood[f]])
            assert np.array_equal(status[f, :ngood[f]], sh) and ninl[f] == nh and np.array_equal(F[f], Fh)
            assert not status[f, ngood[f]:].any()
        prev_kps = kps[n - 1].copy()
    orb.close()
    bf.close()


def test_pipelined_submission_with_filter_equals_blocking_calls(fm):
    """orbx_submit_batch_filtered over 3 batches in flight == extract_batch + match_consecutive + filter_consecutive."""
    w, h, n, nf, nb = 640, 480, 4, 800, 4
    frames = syn.sequence(nb * n, w, h, seed=8)
    orb = ORB(nfeatures=nf, max_size=(w, h), max_batch=n)
    bf = BFMatcher()
    want = []
    for b in range(nb):
        kps, desc, counts = orb.extract_batch(frames[b * n:(b + 1) * n])
        cap = kps.shape[1]
        good, ngood = orb.match_consecutive(bf, 0.8, cap, n)
        want.append((kps, counts, good, ngood) + orb.filter_consecutive(fm, cap, n))
    orb.reset_sequence()
    outs = [(np.zeros((n, cap), KEYPOINT_DTYPE), np.zeros((n, cap, 32), np.uint8), np.zeros(n, np.int32), np.zeros((n, cap), DMATCH_DTYPE),
             np.zeros(n, np.int64), np.zeros((n, cap), np.uint8), np.zeros((n, 3, 3)), np.zeros(n, np.int32)) for _ in range(nb)]
    got = []
    for b in range(nb):
        if orb.batches_in_flight() == orb.pipeline_depth():
            got.append(orb.wait_batch())
        orb.submit_batch(frames[b * n:(b + 1) * n], bf, 0.8, outs[b], fundamental=fm)
    while orb.batches_in_flight():
        got.append(orb.wait_batch())
    assert len(got) == nb
    for b in range(nb):
        kps, counts, good, ngood, status, F, ninl = want[b]
        o = got[b]
        assert np.array_equal(o[2], counts) and np.array_equal(o[4], ngood)
        assert np.array_equal(o[5], status) and np.array_equal(o[6], F) and np.array_equal(o[7], ninl)
        if b:
            assert (ninl >= 0).all()
    orb.close()
    bf.close()


def test_large_pairs_use_the_global_memory_path(fm):
    """More than 9000 correspondences per pair (BASELINE configs[2]: 4K frames, 8000 keypoints, map-sized match lists) do not fit the
    shared-memory copy of the points; the kernel then reads them from global memory -- same results."""
    n = 9600
    p1, p2 = syn.two_view_matches(41, n, 0.6, 0.6, (3840, 2160))
    status, F, ninl = fm.find_batch(p1[None], p2[None], [n])
    Fo, mo, iters = oracle.fm_ransac(p1, p2, 3.0, 0.85)
    assert np.array_equal(status[0], mo) and fm.last_info(1)[0, 1] == iters
    assert rel(F[0], oracle.fm_8point(p1[mo > 0], p2[mo > 0])) <= F_RTOL
    # the same pair inside a batch whose other pairs are small
    q1 = np.zeros((3, n, 2), np.float32)
    q2 = np.zeros((3, n, 2), np.float32)
    q1[1], q2[1] = p1, p2
    q1[0, :500], q2[0, :500] = syn.two_view_matches(42, 500, 0.7, 0.5)
    q1[2, :20], q2[2, :20] = syn.two_view_matches(43, 20, 0.9, 0.3)
    s3, F3, n3 = fm.find_batch(q1, q2, [500, n, 20])
    assert np.array_equal(s3[1], status[0]) and np.array_equal(F3[1], F[0])
    for i, cnt in ((0, 500), (2, 20)):
        _, m_i, _ = oracle.fm_ransac(q1[i, :cnt], q2[i, :cnt], 3.0, 0.85)
        assert np.array_equal(s3[i, :cnt], m_i)


def test_filter_back_matches_per_pair_calls(fm):
    """orbx_filter_back (frame f against its 3 predecessors, history carried across batches) == fmx_compute_fundamental on the
    downloaded keypoints and match lists of every pair."""
    w, h, n, nf, back = 640, 480, 4, 800, 3
    frames = syn.sequence(3 * n, w, h, seed=13)
    orb = ORB(nfeatures=nf, max_size=(w, h), max_batch=n)
    bf = BFMatcher()
    all_kps = []
    for b in range(3):
        kps, desc, counts = orb.extract_batch(frames[b * n:(b + 1) * n])
        cap = kps.shape[1]
        good, ngood = orb.match_back(bf, back, 0.8, cap, n)
        good = good.reshape(n, back, cap)
        ngood = ngood.reshape(n, back)
        status, F, ninl = orb.filter_back(fm, back, cap, n)
        all_kps.extend(kps[i].copy() for i in range(n))
        for f in range(n):
            g = b * n + f
            for j in range(1, back + 1):
                if g - j < 0:
                    assert ninl[f, j - 1] == 0 and not status[f, j - 1].any() and ngood[f, j - 1] == 0
                    continue
                m = good[f, j - 1, :ngood[f, j - 1]]
                Fh, sh, nh = fm.compute_fundamental(all_kps[g], all_kps[g - j], m)
                assert np.array_equal(status[f, j - 1, :len(m)], sh) and ninl[f, j - 1] == nh and np.array_equal(F[f, j - 1], Fh)
    with pytest.raises(Exception):
        orb.extract_batch(frames[:n])
        orb.filter_back(fm, back, cap, n)           # no match_back on the new batch yet
    orb.close()
    bf.close()


@pytest.mark.parametrize("thr,conf,size", [(0.5, 0.85, (1241, 376)), (1.0, 0.99, (1920, 1080)), (3.0, 0.85, (3840, 2160)), (10.0, 0.95, (640, 480))])
def test_fuzz_thresholds_and_image_sizes(fm, thr, conf, size):
    """160 small pairs per configuration (thresholds from 0.5 to 10 px, coordinates up to 4K): every status mask equals the
    oracle's -- the single-precision pre-classification and the division-free test never decide a point differently."""
    npairs, cap = 160, 320
    r = np.random.default_rng(int(thr * 100) + size[0])
    counts = r.integers(15, cap + 1, npairs).astype(np.int32)
    p1 = np.zeros((npairs, cap, 2), np.float32)
    p2 = np.zeros((npairs, cap, 2), np.float32)
    for i in range(npairs):
        a, b = syn.two_view_matches(7000 + 13 * i + size[1], int(counts[i]), float(r.uniform(0.25, 0.98)), float(r.uniform(0.05, 2.0)), size)
        p1[i, :counts[i]], p2[i, :counts[i]] = a, b
    status, F, ninl = fm.find_batch(p1, p2, counts, thr, conf)
    info = fm.last_info(npairs)
    bad = []
    for i in range(npairs):
        n = int(counts[i])
        _, mo, iters = oracle.fm_ransac(p1[i, :n], p2[i, :n], thr, conf)
        if not np.array_equal(status[i, :n], mo) or info[i, 1] != iters:
            bad.append(i)
    assert not bad, "pairs with a different status mask or iteration count: %s" % bad


CHAIN = np.load(os.path.join(os.path.dirname(__file__), "golden", "chain_cases.npz"))


@pytest.mark.parametrize("name", [str(n) for n in CHAIN["names"]])
def test_whole_chain_on_images_matches_cv2(fm, name):
    """The reference's bootstrap sequence on image data, GPU end to end against the cv2 golden (make_golden_chain.py):
    extraction, matchFeatures, computeFundamentalMatrix -- match list and RANSAC status identical, F within tolerance; once
    call by call through host buffers, once in sequence mode with everything resident on the device."""
    seed, w, h, layers, my, mx, nf, ratio = CHAIN[name + "_cfg"]
    w, h, nf, ratio = int(w), int(h), int(nf), float(ratio)
    prev, cur = syn.layered_pair(int(seed), w, h, int(layers), (int(my), int(mx)))
    good = CHAIN[name + "_good"]
    mask = np.unpackbits(CHAIN[name + "_mask"])[:len(good)]
    orb = ORB(nfeatures=nf, max_size=(w, h), max_batch=2)
    bf = BFMatcher()
    kp, dp = orb.detectAndCompute(prev)
    kc, dc = orb.detectAndCompute(cur)
    m = bf.match_ratio(dc, dp, ratio)
    assert list(CHAIN[name + "_counts"][:3]) == [len(kp), len(kc), len(m)]
    assert np.array_equal(m["query_idx"], good[:, 0]) and np.array_equal(m["train_idx"], good[:, 1])
    F, status, ninl = fm.compute_fundamental(kc, kp, m, 3.0, 0.85)
    assert np.array_equal(status, mask) and ninl == CHAIN[name + "_counts"][3]
    assert rel(F, CHAIN[name + "_F8"]) <= F_RTOL
    # sequence mode: both frames as one batch, matches and keypoints never leave the device before the filter
    kps, desc, counts = orb.extract_batch([prev, cur])
    cap = kps.shape[1]
    g, ng = orb.match_consecutive(bf, ratio, cap, 2)
    st, Fs, ni = orb.filter_consecutive(fm, cap, 2, 3.0, 0.85)
    assert ng[1] == len(good) and np.array_equal(st[1, :ng[1]], mask) and np.array_equal(Fs[1], F) and ni[0] == 0
    orb.close()
    bf.close()


def test_pipelined_back_submission_equals_blocking_calls(fm):
    """orbx_submit_batch_back (3 predecessors per frame, matcher + filter, 3 batches in flight) == extract_batch + match_back +
    filter_back per batch, history carried across batches in both."""
    w, h, n, nf, nb, back = 640, 480, 3, 700, 5, 3
    frames = syn.sequence(nb * n, w, h, seed=17)
    orb = ORB(nfeatures=nf, max_size=(w, h), max_batch=n)
    bf = BFMatcher()
    want = []
    for b in range(nb):
        kps, desc, counts = orb.extract_batch(frames[b * n:(b + 1) * n])
        cap = kps.shape[1]
        good, ngood = orb.match_back(bf, back, 0.8, cap, n)
        want.append((counts.copy(), good.copy(), ngood.copy()) + orb.filter_back(fm, back, cap, n))
    orb.reset_sequence()

    def bufs():
        return (np.zeros((n, cap), KEYPOINT_DTYPE), np.zeros((n, cap, 32), np.uint8), np.zeros(n, np.int32), np.zeros((n, back, cap), DMATCH_DTYPE),
                np.zeros((n, back), np.int64), np.zeros((n, back, cap), np.uint8), np.zeros((n, back, 3, 3)), np.zeros((n, back), np.int32))
    got = []
    for b in range(nb):
        if orb.batches_in_flight() == orb.pipeline_depth():
            got.append(orb.wait_batch())
        orb.submit_batch(frames[b * n:(b + 1) * n], bf, 0.8, bufs(), fundamental=fm, back=back)
    while orb.batches_in_flight():
        got.append(orb.wait_batch())
    for b in range(nb):
        counts, good, ngood, status, F, ninl = want[b]
        o = got[b]
        assert np.array_equal(o[2], counts) and np.array_equal(o[4], ngood)
        for f in range(n):
            for j in range(back):
                k = ngood[f, j]
                assert np.array_equal(o[3][f, j, :k], good[f, j, :k]) and np.array_equal(o[5][f, j, :k], status[f, j, :k])
        assert np.array_equal(o[6], F) and np.array_equal(o[7], ninl)
    assert want[-1][2].min() > 100 and want[-1][5].min() >= 0
    # matcher only (no filter) through the same entry point
    orb.reset_sequence()
    o = bufs()[:5]
    orb.submit_batch(frames[:n], bf, 0.8, o, back=back)
    orb.wait_batch()
    assert np.array_equal(o[4], want[0][2])
    orb.close()
    bf.close()


def test_regression_two_improvements_in_one_iteration(fm):
    """Found by tests/soak_fmat.py: iteration 12 of this pair yields candidates with 10, 67 and 71 inliers.  The 67 lowers the budget
    to 8 (< 12), but OpenCV checks the budget once per iteration, so the 71 of the same iteration is still scored and wins.
    (cv2 mask in tests/golden/fmat_regressions.npz.)"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fmat_regressions.npz"))
    p1, p2 = g["two_improvements_in_one_iteration_p1"], g["two_improvements_in_one_iteration_p2"]
    thr, conf = g["two_improvements_in_one_iteration_cfg"]
    status, F, ninl = fm.find_batch(p1[None], p2[None], [len(p1)], float(thr), float(conf))
    assert np.array_equal(status[0], g["two_improvements_in_one_iteration_mask"]) and ninl[0] == 71
    assert fm.last_info(1)[0, 1] == 13


def test_least_median_path_for_8_to_14_points(fm):
    """8..14 correspondences: OpenCV switches from RANSAC to LMedS.  With 14 points the result is reproducible and must equal the
    oracle's (which equals cv2's, tests/golden n14_lmeds_*); with fewer the winner is decided by rounding noise even inside OpenCV,
    so only the contract is checked: either no model, or >= 7 inliers all within the reported threshold of a rank-2 F."""
    npairs = 40
    p1 = np.zeros((npairs, 14, 2), np.float32)
    p2 = np.zeros((npairs, 14, 2), np.float32)
    for i in range(npairs):
        p1[i], p2[i] = syn.two_view_matches(3000 + i, 14, 0.6 + 0.01 * i, 0.3 + 0.02 * i)
    for conf in (0.85, 0.99):
        status, F, ninl = fm.find_batch(p1, p2, np.full(npairs, 14, np.int32), 3.0, conf)
        info = fm.last_info(npairs)
        for i in range(npairs):
            Fo, mo, iters = oracle.fm_ransac(p1[i], p2[i], 3.0, conf)
            assert np.array_equal(status[i], mo) and ninl[i] == mo.sum() and info[i, 1] == iters, i
            if mo.sum() >= 8:
                assert rel(F[i], oracle.fm_8point(p1[i][mo > 0], p2[i][mo > 0])) <= F_RTOL
    counts = np.array([8 + i % 6 for i in range(npairs)], np.int32)      # 8..13
    status, F, ninl = fm.find_batch(p1, p2, counts, 3.0, 0.85)
    for i in range(npairs):
        assert not status[i, counts[i]:].any() and ninl[i] == status[i].sum() and (ninl[i] == 0 or ninl[i] >= 7)
        if ninl[i] >= 8:
            assert abs(np.linalg.det(F[i] / np.linalg.norm(F[i]))) < 1e-9


@pytest.mark.parametrize("conf", [1e-9, 1e-4, 0.3, 1.0 - 1e-12])
def test_extreme_confidences(fm, conf):
    """Confidence only enters through the iteration budget log(1 - conf) / log(1 - w^7): tiny values end the loop after a handful of
    iterations, values next to 1 run it to the cap.  The budget screen in single precision must never change the count."""
    npairs, cap = 24, 400
    r = np.random.default_rng(11)
    counts = r.integers(15, cap + 1, npairs).astype(np.int32)
    p1 = np.zeros((npairs, cap, 2), np.float32)
    p2 = np.zeros((npairs, cap, 2), np.float32)
    for i in range(npairs):
        p1[i, :counts[i]], p2[i, :counts[i]] = syn.two_view_matches(8100 + i, int(counts[i]), float(r.uniform(0.1, 0.95)), 0.5)
    status, F, ninl = fm.find_batch(p1, p2, counts, 3.0, conf)
    info = fm.last_info(npairs)
    for i in range(npairs):
        n = int(counts[i])
        _, mo, iters = oracle.fm_ransac(p1[i, :n], p2[i, :n], 3.0, conf)
        assert info[i, 1] == iters and np.array_equal(status[i, :n], mo), (i, info[i, 1], iters)


@pytest.mark.parametrize("step", [8, 32, 64])
def test_collinear_rejections_on_grid_points(fm, step):
    """Positions snapped to a coarse grid make 3 % (step 8) to 50 % (step 64) of the drawn samples fail OpenCV's collinearity test.
    The kernel draws a round's samples assuming none is rejected and redoes the round from the first rejected one with the
    sequential rule; the generator state must continue exactly as in the sequential loop -- same status masks and iteration
    counts as the oracle."""
    npairs, cap = 32, 400
    r = np.random.default_rng(step)
    counts = r.integers(60, cap + 1, npairs).astype(np.int32)
    p1 = np.zeros((npairs, cap, 2), np.float32)
    p2 = np.zeros((npairs, cap, 2), np.float32)
    for i in range(npairs):
        a, b = syn.two_view_matches(9100 + i, int(counts[i]), float(r.uniform(0.3, 0.9)), 0.3, (640, 480))
        p1[i, :counts[i]], p2[i, :counts[i]] = np.round(a / step) * step, np.round(b / step) * step
    thr = max(3.0, step / 2)
    status, F, ninl = fm.find_batch(p1, p2, counts, thr, 0.95)
    info = fm.last_info(npairs)
    for i in range(npairs):
        n = int(counts[i])
        _, mo, iters = oracle.fm_ransac(p1[i, :n], p2[i, :n], thr, 0.95)
        assert info[i, 1] == iters and np.array_equal(status[i, :n], mo), (i, int(info[i, 1]), iters)
