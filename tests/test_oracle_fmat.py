"""CPU: the fundamental-matrix oracle (oracle/fmat_oracle.c) against the cv2 4.13.0 golden vectors
(tests/golden/fmat_cases.npz, made by tests/golden/make_golden_fmat.py) and known answers.

Tolerances: RANSAC status masks are index work -- identical.  Matrices are double-precision linear algebra done with a
different (but equally valid) decomposition than OpenCV's: relative Frobenius error <= 1e-6 after scaling to F[2,2] = 1
(measured: ~1e-12 for the RANSAC candidate, ~1e-10 for the 8-point matrix)."""
import os

import numpy as np
import pytest

import oracle
from monocular_slam_b200 import synthetic as syn

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "fmat_cases.npz"))
NAMES = [str(n) for n in GOLD["names"]]
F_RTOL = 1e-6


def case(name):
    seed, n, inl, noise, thr, conf = GOLD[name + "_cfg"]
    p1, p2 = syn.two_view_matches(int(seed), int(n), float(inl), float(noise))
    mask = np.unpackbits(GOLD[name + "_mask"])[:int(n)]
    return p1, p2, float(thr), float(conf), mask, GOLD[name + "_Fransac"], GOLD[name + "_F8"]


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("name", NAMES)
def test_ransac_mask_and_matrices_match_cv2(name):
    p1, p2, thr, conf, mask, Fr, F8 = case(name)
    F, m, iters = oracle.fm_ransac(p1, p2, thr, conf)
    assert F is not None and 1 <= iters <= 1000
    assert np.array_equal(m, mask), "%s: %d status bytes differ" % (name, int((m != mask).sum()))
    assert rel(F, Fr) <= F_RTOL
    F8o = oracle.fm_8point(p1[mask > 0], p2[mask > 0])
    assert rel(F8o, F8) <= F_RTOL


@pytest.mark.parametrize("name", NAMES[:4])
def test_compute_fundamental_is_ransac_then_8point(name):
    p1, p2, thr, conf, mask, _, F8 = case(name)
    n = len(p1)
    perm = np.random.default_rng(1).permutation(n)
    # keypoint tables in a different order than the match list, as in the reference (positions1[matches[i].queryIdx])
    pos1 = np.zeros((n, 2), np.float32)
    pos1[perm] = p1
    matches = np.zeros((n, 4), np.int32)            # cv::DMatch rows: queryIdx, trainIdx, imgIdx, distance
    matches[:, 0] = perm
    matches[:, 1] = np.arange(n)
    F, status, ninl = oracle.compute_fundamental(pos1, p2, matches, thr, conf)
    assert np.array_equal(status, mask) and ninl == int(mask.sum())
    assert rel(F, F8) <= F_RTOL


def test_solve_cubic_known_answers():
    assert np.allclose(oracle.solve_cubic([1, -6, 11, -6]), [1, 3, 2])          # OpenCV's root order
    assert np.allclose(oracle.solve_cubic([2, 1, 3, -6]), [1])
    assert np.allclose(oracle.solve_cubic([0, 1, -3, 2]), [2, 1])
    assert np.allclose(oracle.solve_cubic([0, 0, 2, -4]), [2])
    assert len(oracle.solve_cubic([0, 0, 0, 1])) == 0


def test_seven_point_candidates_satisfy_their_sample():
    p1, p2 = syn.two_view_matches(3, 7, 1.0, 0.0)
    Fs = oracle.fm_7point(p1, p2)
    assert 1 <= len(Fs) <= 3
    h1 = np.c_[p1.astype(np.float64), np.ones(7)]
    h2 = np.c_[p2.astype(np.float64), np.ones(7)]
    for F in Fs:
        assert abs(np.linalg.det(F / np.linalg.norm(F))) < 1e-12
        assert np.abs(np.einsum("ij,jk,ik->i", h2, F, h1)).max() / np.linalg.norm(F) < 1e-6
        assert oracle.fm_errors(p1, p2, F).max() < 1e-10


def test_eight_point_recovers_exact_geometry():
    p1, p2 = syn.two_view_matches(5, 200, 1.0, 0.0)
    F = oracle.fm_8point(p1, p2)
    assert abs(np.linalg.det(F / np.linalg.norm(F))) < 1e-12
    assert oracle.fm_errors(p1, p2, F).max() < 1e-3         # float32 positions of exact projections
    assert oracle.fm_8point(p1[:7], p2[:7]) is None


def test_all_outliers_and_small_sets():
    r = np.random.default_rng(0)
    p1 = r.uniform(0, 1000, (200, 2)).astype(np.float32)
    p2 = r.uniform(0, 1000, (200, 2)).astype(np.float32)
    F, m, iters = oracle.fm_ransac(p1, p2, 3.0, 0.85)
    assert iters == 1000 and (F is None or m.sum() >= 7)
    F, status, ninl = oracle.compute_fundamental(p1[:5], p2[:5], np.zeros((5, 4), np.int32), 3.0, 0.85)
    assert ninl == 0 and not F.any()


CHAIN = np.load(os.path.join(os.path.dirname(__file__), "golden", "chain_cases.npz"))


@pytest.mark.parametrize("name", [str(n) for n in CHAIN["names"]])
def test_whole_chain_on_images_matches_cv2(name):
    """extract -> matchFeatures -> computeFundamentalMatrix on two views of a layered-depth scene, oracle vs the cv2 golden
    (tests/golden/make_golden_chain.py): same match list, same RANSAC status, same 8-point matrix."""
    seed, w, h, layers, my, mx, nf, ratio = CHAIN[name + "_cfg"]
    prev, cur = syn.layered_pair(int(seed), int(w), int(h), int(layers), (int(my), int(mx)))
    P = oracle.Params(nfeatures=int(nf))
    kp, dp = oracle.detect_and_compute(prev, P)
    kc, dc = oracle.detect_and_compute(cur, P)
    gq, gt, gd = oracle.match_features(dc, dp, float(ratio))
    good = CHAIN[name + "_good"]
    assert list(CHAIN[name + "_counts"][:3]) == [len(kp), len(kc), len(gq)]
    assert np.array_equal(gq, good[:, 0]) and np.array_equal(gt, good[:, 1]) and np.array_equal(gd, good[:, 2])
    matches = np.zeros((len(gq), 4), np.int32)
    matches[:, 0], matches[:, 1] = gq, gt
    F, status, ninl = oracle.compute_fundamental(np.c_[kc["x"], kc["y"]], np.c_[kp["x"], kp["y"]], matches, 3.0, 0.85)
    mask = np.unpackbits(CHAIN[name + "_mask"])[:len(gq)]
    assert np.array_equal(status, mask) and ninl == CHAIN[name + "_counts"][3]
    assert rel(F, CHAIN[name + "_F8"]) <= F_RTOL


def test_regression_two_improvements_in_one_iteration():
    """The budget is checked once per iteration: a second improvement inside the iteration that exhausted it still counts."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fmat_regressions.npz"))
    p1, p2 = g["two_improvements_in_one_iteration_p1"], g["two_improvements_in_one_iteration_p2"]
    thr, conf = g["two_improvements_in_one_iteration_cfg"]
    F, m, iters = oracle.fm_ransac(p1, p2, float(thr), float(conf))
    assert np.array_equal(m, g["two_improvements_in_one_iteration_mask"]) and iters == 13
