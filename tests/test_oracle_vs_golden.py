"""CPU: the oracle (oracle/orb_oracle.c) must reproduce every golden vector produced by cv2 (tests/golden/make_golden.py).

This is what pins the oracle: the reference has no fixtures of its own for this path."""
import os

import numpy as np
import pytest

import oracle
from helpers import assert_descriptors_equal, assert_keypoints_equal, sha
from monocular_slam_b200 import synthetic as syn


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("score", ["harris", "fast"])
def test_kitti_pair_extraction(golden_dir, score):
    g = _load(golden_dir, "kitti_pair.npz")
    big = g["canvas"]
    frames = [np.ascontiguousarray(big[:376, :1241]), np.ascontiguousarray(big[3:, 7:])]
    P = oracle.Params(nfeatures=2000, score_type=0 if score == "harris" else 1)
    for i, f in enumerate(frames):
        k, d = oracle.detect_and_compute(f, P)
        assert_keypoints_equal(k, g[f"{score}_kp{i}"], f"{score} frame {i}")
        assert_descriptors_equal(d, g[f"{score}_desc{i}"], f"{score} frame {i}")


@pytest.mark.parametrize("score", ["harris", "fast"])
def test_kitti_pair_matching(golden_dir, score):
    g = _load(golden_dir, "kitti_pair.npz")
    idx, dist = oracle.knn2(g[f"{score}_desc0"], g[f"{score}_desc1"])
    assert np.array_equal(idx, g[f"{score}_knn_idx"]) and np.array_equal(dist, g[f"{score}_knn_dist"])
    for r in (0.75, 0.8, 0.85):
        q, t, d = oracle.ratio_test(idx, dist, r)
        assert np.array_equal(np.stack([q, t, d], 1), g[f"{score}_good_{int(r * 100)}"])


def test_pyramid_hashes(golden_dir):
    g = _load(golden_dir, "kitti_pair.npz")
    f0 = np.ascontiguousarray(g["canvas"][:376, :1241])
    P = oracle.Params(nfeatures=2000)
    for l in range(8):
        assert sha(oracle.pyramid_level(P, f0, l)) == str(g["pyr_sha_f0"][l]), "level %d" % l


def test_hd_frame(golden_dir):
    g = _load(golden_dir, "hd_frames.npz")
    img = syn.frame(1, 1920, 1080)
    if sha(img) != str(g["s1_img_sha"]):
        pytest.skip("numpy on this host generates a different synthetic frame than the one the golden was made from")
    k, d = oracle.detect_and_compute(img, oracle.Params(nfeatures=2000))
    assert_keypoints_equal(k, g["s1_kp"], "hd s1")
    assert_descriptors_equal(d, g["s1_desc"], "hd s1")
    idx, dist = oracle.knn2(g["s1_desc"], g["s2_desc"])
    assert np.array_equal(idx, g["knn_idx"]) and np.array_equal(dist, g["knn_dist"])


SMALL = ["tiny_97x71", "small_200x150", "odd_333x257", "thin_300x63", "flat_400x300", "checker_640x480", "textured_640x480"]


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("score", ["harris", "fast"])
def test_small_frames(golden_dir, name, score):
    g = _load(golden_dir, "small_frames.npz")
    k, d = oracle.detect_and_compute(g[f"{name}_img"], oracle.Params(nfeatures=500, score_type=0 if score == "harris" else 1))
    assert_keypoints_equal(k, g[f"{name}_{score}_kp"], name)
    assert_descriptors_equal(d, g[f"{name}_{score}_desc"], name)


def test_compute_border_keypoints(golden_dir):
    g = _load(golden_dir, "small_frames.npz")
    k, d = oracle.compute(g["border_img"], g["border_kp_in"], oracle.Params(nfeatures=700))
    assert_keypoints_equal(k, g["border_kp_out"], "border")
    assert_descriptors_equal(d, g["border_desc"], "border")


MATCH = ["planted", "dup_rows", "zeros", "nt1", "nt2", "low_entropy", "ragged_33x65"]


@pytest.mark.parametrize("name", MATCH)
def test_matcher_cases(golden_dir, name):
    g = _load(golden_dir, "matcher_cases.npz")
    idx, dist = oracle.knn2(g[f"{name}_q"], g[f"{name}_t"])
    assert np.array_equal(idx, g[f"{name}_idx"]) and np.array_equal(dist, g[f"{name}_dist"])
    if len(g[f"{name}_t"]) >= 2:
        for r in (0.75, 0.8, 0.85):
            q, t, d = oracle.ratio_test(idx, dist, r)
            assert np.array_equal(np.stack([q, t, d], 1).reshape(-1, 3), g[f"{name}_good_{int(r * 100)}"])


def test_empty_inputs():
    idx, dist = oracle.knn2(np.zeros((0, 32), np.uint8), syn.descriptors(1, 5))
    assert idx.shape == (0, 2)
    idx, dist = oracle.knn2(syn.descriptors(1, 5), np.zeros((0, 32), np.uint8))
    assert (idx == -1).all() and (dist == -1).all()


def test_gauss_kernel_constants():
    want = [float.fromhex(v) for v in ("0x1.1f5f62p-4", "0x1.0c70fcp-3", "0x1.869472p-3", "0x1.ba95c0p-3")]
    k = oracle.gauss_kernel7()
    assert [float(x) for x in k] == want + want[2::-1]


def test_quotas_and_sizes():
    P = oracle.Params(nfeatures=2000)
    assert oracle.level_quotas(P) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert oracle.level_sizes(P, 1920, 1080) == [(1920, 1080), (1600, 900), (1333, 750), (1111, 625), (926, 521), (772, 434),
                                                 (643, 362), (536, 301)]
    assert oracle.level_quotas(oracle.Params(nfeatures=500)) == [109, 90, 75, 63, 52, 44, 36, 31]


def test_bgr_input(golden_dir):
    """3-channel frames: the oracle's cvtColor(BGR2GRAY) restatement and the extraction behind it reproduce cv2."""
    g = np.load(os.path.join(golden_dir, "bgr_frame.npz"))
    assert np.array_equal(oracle.bgr2gray(g["lattice"]), g["lattice_gray"])
    img = g["img"]
    assert sha(oracle.bgr2gray(img)) == str(g["gray_sha"])
    k, d = oracle.detect_and_compute(img, oracle.Params(nfeatures=500))
    assert_keypoints_equal(k, g["kp"], "bgr 333x257")
    assert_descriptors_equal(d, g["desc"], "bgr 333x257")
    big = syn.bgr_frame(12, 640, 480)
    assert sha(big) == str(g["big_sha"]) and sha(oracle.bgr2gray(big)) == str(g["big_gray_sha"])
    k, d = oracle.detect_and_compute(big, oracle.Params(nfeatures=1000))
    assert_keypoints_equal(k, g["big_kp"], "bgr 640x480")
    assert_descriptors_equal(d, g["big_desc"], "bgr 640x480")


PARAM_CASES = ["l4_s15", "l1", "l12_s11_t30", "t10_fast", "l3_s20_t5", "l16_s105"]


@pytest.mark.parametrize("name", PARAM_CASES)
def test_parameter_surface(golden_dir, name):
    """nlevels / scaleFactor / fastThreshold / scoreType / nfeatures other than the defaults (cv2 golden)."""
    g = np.load(os.path.join(golden_dir, "params_cases.npz"))
    img = syn.frame(9, 640, 480)
    assert sha(img) == str(g["img_sha"])
    nf, sf, nl, st, thr = g[name + "_params"]
    P = oracle.Params(nfeatures=int(nf), scale_factor=float(sf), nlevels=int(nl), score_type=int(st), fast_threshold=int(thr))
    k, d = oracle.detect_and_compute(img, P)
    assert_keypoints_equal(k, g[name + "_kp"], name)
    assert_descriptors_equal(d, g[name + "_desc"], name)
