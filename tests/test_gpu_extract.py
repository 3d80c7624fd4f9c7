"""GPU parity: orbx_* (K1-K5) through the C ABI against the oracle and the cv2 golden vectors.

Bar (BASELINE.json north_star): keypoint coordinates, scores/responses, octaves bit-exact; angles bit-exact here
(tolerance allowed: 1e-4 rad); descriptor bytes bit-exact (allowance: 0.1% of descriptors, reported)."""
import os

import numpy as np
import pytest

import oracle
from helpers import assert_descriptors_equal, assert_keypoints_equal, sha
from monocular_slam_b200 import FAST_SCORE, HARRIS_SCORE, ORB, DataManager, FeatureExtractor, OrbxError
from monocular_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _orb(nf, score, size, batch=1):
    return ORB(nfeatures=nf, scoreType=score, max_size=size, max_batch=batch)


@pytest.mark.parametrize("w,h", [(1241, 376), (640, 480), (333, 257), (1920, 1080)])
def test_pyramid_levels(w, h):
    img = syn.frame(w + h, w, h)
    orb = _orb(500, HARRIS_SCORE, (w, h))
    P = oracle.Params(nfeatures=500)
    for l in range(8):
        got = orb.debug_pyramid_level(img, l)
        want = oracle.pyramid_level(P, img, l)
        assert got.shape == want.shape and np.array_equal(got, want), "level %d: %d px differ" % (l, (got != want).sum())
    orb.close()


def test_pyramid_golden_hashes(golden_dir):
    g = np.load(os.path.join(golden_dir, "kitti_pair.npz"))
    f0 = np.ascontiguousarray(g["canvas"][:376, :1241])
    orb = _orb(2000, HARRIS_SCORE, (1241, 376))
    for l in range(8):
        assert sha(orb.debug_pyramid_level(f0, l)) == str(g["pyr_sha_f0"][l]), "level %d" % l
    orb.close()


@pytest.mark.parametrize("w,h", [(1241, 376), (640, 480), (200, 150)])
def test_fast_levels(w, h):
    img = syn.frame(3 * w + h, w, h)
    orb = _orb(500, HARRIS_SCORE, (w, h))
    P = oracle.Params(nfeatures=500)
    for l in range(8):
        xs, ys, sc = orb.debug_fast_level(img, l)
        ox, oy, osc = oracle.level_fast(img, P, l)
        assert len(xs) == len(ox), "level %d: %d corners, expected %d" % (l, len(xs), len(ox))
        assert np.array_equal(xs, ox) and np.array_equal(ys, oy) and np.array_equal(sc, osc), "level %d" % l
    orb.close()


@pytest.mark.parametrize("score", ["harris", "fast"])
def test_kitti_pair_golden(golden_dir, score):
    g = np.load(os.path.join(golden_dir, "kitti_pair.npz"))
    big = g["canvas"]
    frames = [np.ascontiguousarray(big[:376, :1241]), np.ascontiguousarray(big[3:, 7:])]
    orb = _orb(2000, HARRIS_SCORE if score == "harris" else FAST_SCORE, (1241, 376))
    for i, f in enumerate(frames):
        # the reference's call order: detect, then compute (src/FeatureExtractor.cpp:17,19)
        k = orb.detect(f)
        assert_keypoints_equal(k, g[f"{score}_kp{i}"], "detect %s %d" % (score, i))
        k2, d = orb.compute(f, k)
        assert_keypoints_equal(k2, g[f"{score}_kp{i}"], "compute %s %d" % (score, i))
        assert_descriptors_equal(d, g[f"{score}_desc{i}"], "compute %s %d" % (score, i))
        k3, d3 = orb.detectAndCompute(f)
        assert_keypoints_equal(k3, k2, "detectAndCompute")
        assert_descriptors_equal(d3, d, "detectAndCompute")
    orb.close()


def test_hd_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "hd_frames.npz"))
    orb = _orb(2000, HARRIS_SCORE, (1920, 1080))
    for seed in (1, 2):
        img = syn.frame(seed, 1920, 1080)
        k, d = orb.detectAndCompute(img)
        if sha(img) == str(g[f"s{seed}_img_sha"]):
            assert_keypoints_equal(k, g[f"s{seed}_kp"], "hd golden %d" % seed)
            assert_descriptors_equal(d, g[f"s{seed}_desc"], "hd golden %d" % seed)
        ok, od = oracle.detect_and_compute(img, oracle.Params(nfeatures=2000))
        assert_keypoints_equal(k, ok, "hd oracle %d" % seed)
        assert_descriptors_equal(d, od, "hd oracle %d" % seed)
    orb.close()


SMALL = ["tiny_97x71", "small_200x150", "odd_333x257", "thin_300x63", "flat_400x300", "checker_640x480", "textured_640x480"]


@pytest.mark.parametrize("score", ["harris", "fast"])
def test_small_frames_golden(golden_dir, score):
    g = np.load(os.path.join(golden_dir, "small_frames.npz"))
    orb = _orb(500, HARRIS_SCORE if score == "harris" else FAST_SCORE, (640, 480))
    for name in SMALL:
        k, d = orb.detectAndCompute(g[f"{name}_img"])
        assert_keypoints_equal(k, g[f"{name}_{score}_kp"], name)
        assert_descriptors_equal(d, g[f"{name}_{score}_desc"], name)
    orb.close()


def test_compute_border_keypoints_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "small_frames.npz"))
    orb = _orb(700, HARRIS_SCORE, (800, 600))
    k, d = orb.compute(g["border_img"], g["border_kp_in"])
    assert_keypoints_equal(k, g["border_kp_out"], "border")
    assert_descriptors_equal(d, g["border_desc"], "border")
    orb.close()


def test_4k_8000(golden_dir):
    img = syn.frame(11, 3840, 2160)
    orb = _orb(8000, HARRIS_SCORE, (3840, 2160))
    k, d = orb.detectAndCompute(img)
    ok, od = oracle.detect_and_compute(img, oracle.Params(nfeatures=8000))
    assert_keypoints_equal(k, ok, "4k")
    assert_descriptors_equal(d, od, "4k")
    orb.close()


def test_batch_equals_single():
    seq = syn.sequence(6, 1241, 376, seed=5)
    orb = _orb(2000, HARRIS_SCORE, (1241, 376), batch=6)
    kps, desc, counts = orb.extract_batch(list(seq))
    P = oracle.Params(nfeatures=2000)
    for i in range(len(seq)):
        ok, od = oracle.detect_and_compute(seq[i], P)
        assert counts[i] == len(ok)
        assert_keypoints_equal(kps[i, :counts[i]], ok, "batch frame %d" % i)
        assert_descriptors_equal(desc[i, :counts[i]], od, "batch frame %d" % i)
    # a strided view (row stride > width) must give the same result
    wide = np.zeros((376, 1300), np.uint8)
    wide[:, :1241] = seq[0]
    k, d = orb.detectAndCompute(wide[:, :1241])
    assert_keypoints_equal(k, kps[0, :counts[0]], "strided")
    orb.close()


def test_handle_reuse_across_sizes():
    orb = _orb(500, HARRIS_SCORE, (800, 600))
    P = oracle.Params(nfeatures=500)
    for (w, h) in [(800, 600), (320, 240), (799, 451), (800, 600)]:
        img = syn.frame(w, w, h)
        k, d = orb.detectAndCompute(img)
        ok, od = oracle.detect_and_compute(img, P)
        assert_keypoints_equal(k, ok, "%dx%d" % (w, h))
        assert_descriptors_equal(d, od, "%dx%d" % (w, h))
    with pytest.raises(OrbxError):
        orb.detect(syn.frame(1, 801, 600))
    orb.close()


def test_capacity_error_and_retry():
    img = syn.frame(1, 640, 480)
    orb = _orb(500, FAST_SCORE, (640, 480))
    n = len(oracle.detect(img, oracle.Params(nfeatures=500, score_type=1)))
    assert n > 500    # FAST_SCORE keeps ties
    k = orb.detect(img, cap=n)
    assert len(k) == n
    k = orb.detect(img, cap=500)     # too small: the wrapper retries with the handle's maximum
    assert len(k) == n
    orb.close()


def test_feature_extractor_node():
    """FeatureExtractor::process fills Features exactly as src/FeatureExtractor.cpp:13-31 does."""
    seq = syn.sequence(2, 640, 480, seed=9)
    dm = DataManager(list(seq))
    node = FeatureExtractor(nfeatures=500, max_size=(640, 480))
    node.init()
    for i in range(2):
        assert node.validationCheck(dm, i)
        node.process(dm, i)
        ok, od = oracle.detect_and_compute(seq[i], oracle.Params(nfeatures=500))
        f = dm.frames[i].features
        assert np.array_equal(f.descriptors, od)
        assert np.array_equal(f.positions, np.stack([ok["x"], ok["y"]], 1).astype(np.float64))
        assert np.array_equal(f.scales, ok["size"].astype(np.float64))
        assert (f.mapPointsIndices == -1).all() and len(f.mapPointsIndices) == len(ok)
    node.destroy()


def test_bad_arguments():
    with pytest.raises(OrbxError):
        ORB(nfeatures=500, edgeThreshold=19)
    with pytest.raises(OrbxError):
        ORB(nfeatures=-1)
    orb = _orb(500, HARRIS_SCORE, (64, 64))
    with pytest.raises(ValueError):
        orb.detect(np.zeros((10, 10, 4), np.uint8))          # only CV_8UC1 and CV_8UC3 exist on this path
    with pytest.raises(ValueError):
        orb.detect(np.zeros((10, 10), np.float32))
    assert len(orb.detect(np.zeros((10, 10, 3), np.uint8))) == 0   # 3-channel frames are accepted (converted to gray)
    k = orb.detect(np.zeros((64, 64), np.uint8))
    assert len(k) == 0
    orb.close()


def test_bgr_input_golden(golden_dir):
    """CV_8UC3 frames: gray conversion on the device (cvtColor BGR2GRAY arithmetic) + extraction == cv2 on the BGR frame;
    single-frame, batched, pipelined and device-pointer entry points, odd width (unaligned rows) and aligned width."""
    import torch
    g = np.load(os.path.join(golden_dir, "bgr_frame.npz"))
    img = g["img"]
    orb = ORB(nfeatures=500, max_size=(640, 480), max_batch=2)
    k = orb.detect(img)
    k, d = orb.compute(img, k)
    assert_keypoints_equal(k, g["kp"], "bgr detect+compute")
    assert_descriptors_equal(d, g["desc"], "bgr detect+compute")
    gray = oracle.bgr2gray(img)
    for level in (0, 3):
        assert np.array_equal(orb.debug_pyramid_level(img, level), orb.debug_pyramid_level(gray, level))
    kps, desc, counts = orb.extract_batch([img, img])
    for i in range(2):
        assert_keypoints_equal(kps[i, :counts[i]], g["kp"], "bgr batch %d" % i)
        assert_descriptors_equal(desc[i, :counts[i]], g["desc"], "bgr batch %d" % i)
    orb.close()
    big = syn.bgr_frame(12, 640, 480)
    orb = ORB(nfeatures=1000, max_size=(640, 480), max_batch=2)
    kps, desc, counts = orb.extract_batch([big, big])
    assert_keypoints_equal(kps[1, :counts[1]], g["big_kp"], "bgr 640x480 batch")
    assert_descriptors_equal(desc[1, :counts[1]], g["big_desc"], "bgr 640x480 batch")
    # device-resident BGR frames
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        orb.set_stream(stream.cuda_stream)
        orb.set_input_channels(3)
        cap = orb.default_cap
        d_frames = torch.from_numpy(np.stack([big, big])).cuda()
        d_kps = torch.empty((2, cap, 7), dtype=torch.float32, device="cuda")
        d_desc = torch.empty((2, cap, 32), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
        orb.extract_batch_dev(d_frames.data_ptr(), 640 * 480 * 3, 2, 640, 480, 640 * 3, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_cnt.data_ptr())
        orb.check_dev()
        stream.synchronize()
    from monocular_slam_b200 import KEYPOINT_DTYPE
    n = int(d_cnt[0])
    assert_keypoints_equal(d_kps.cpu().numpy().view(KEYPOINT_DTYPE).reshape(2, cap)[0, :n], g["big_kp"], "bgr dev")
    assert_descriptors_equal(d_desc.cpu().numpy()[0, :n], g["big_desc"], "bgr dev")
    orb.set_stream(0)
    # back to gray on the same handle
    k1, d1 = orb.detectAndCompute(oracle.bgr2gray(big))
    assert_keypoints_equal(k1, g["big_kp"], "gray after bgr")
    orb.close()


@pytest.mark.parametrize("name", ["l4_s15", "l1", "l12_s11_t30", "t10_fast", "l3_s20_t5", "l16_s105"])
def test_parameter_surface_golden(golden_dir, name):
    """Every runtime parameter orbx_create accepts, away from its default, against cv2's output (params_cases.npz)."""
    g = np.load(os.path.join(golden_dir, "params_cases.npz"))
    img = syn.frame(9, 640, 480)
    nf, sf, nl, st, thr = g[name + "_params"]
    orb = ORB(nfeatures=int(nf), scaleFactor=float(sf), nlevels=int(nl), scoreType=int(st), fastThreshold=int(thr), max_size=(640, 480),
              max_batch=2)
    k, d = orb.detectAndCompute(img)
    assert_keypoints_equal(k, g[name + "_kp"], name)
    assert_descriptors_equal(d, g[name + "_desc"], name)
    k2 = orb.detect(img)
    k2, d2 = orb.compute(img, k2)
    assert_keypoints_equal(k2, g[name + "_kp"], name + " detect->compute")
    assert_descriptors_equal(d2, g[name + "_desc"], name + " detect->compute")
    kps, desc, counts = orb.extract_batch([img, img])
    assert_keypoints_equal(kps[1, :counts[1]], g[name + "_kp"], name + " batch")
    assert_descriptors_equal(desc[1, :counts[1]], g[name + "_desc"], name + " batch")
    orb.close()


FUZZ = [  # (w, h, nfeatures, scaleFactor, nlevels, scoreType, fastThreshold, seed)
    (67, 64, 50, 1.2, 8, 0, 20, 1), (64, 67, 50, 1.2, 3, 1, 5, 2), (130, 95, 200, 1.3, 6, 0, 12, 3), (257, 129, 300, 1.25, 5, 1, 20, 4),
    (511, 300, 900, 1.2, 8, 0, 30, 5), (640, 65, 120, 1.15, 4, 0, 20, 6), (63, 400, 100, 1.2, 8, 0, 20, 7), (1023, 257, 1500, 1.2, 8, 1, 8, 8),
    (385, 385, 4000, 1.1, 10, 0, 3, 9), (1280, 720, 1000, 1.41, 6, 0, 20, 10), (129, 128, 500, 2.0, 4, 0, 10, 11), (800, 608, 250, 1.2, 1, 1, 40, 12),
]


@pytest.mark.parametrize("case", FUZZ, ids=lambda c: "%dx%d_nf%d_s%g_l%d_t%d_%d" % (c[0], c[1], c[2], c[3], c[4], c[6], c[5]))
def test_fuzz_against_oracle(case):
    """Odd sizes (down to frames with no interior at the coarse levels), every runtime parameter, both score types,
    dense and sparse thresholds: detectAndCompute == oracle, and detect -> compute == detectAndCompute."""
    w, h, nf, sf, nl, st, thr, seed = case
    img = syn.textured_frame(seed, w, h) if seed % 2 else syn.frame(seed, w, h, nrect=max(10, w * h // 4000))
    P = oracle.Params(nfeatures=nf, scale_factor=sf, nlevels=nl, score_type=st, fast_threshold=thr)
    ok, od = oracle.detect_and_compute(img, P)
    orb = ORB(nfeatures=nf, scaleFactor=sf, nlevels=nl, scoreType=st, fastThreshold=thr, max_size=(w, h), max_batch=2)
    k, d = orb.detectAndCompute(img)
    assert_keypoints_equal(k, ok, "detectAndCompute")
    assert_descriptors_equal(d, od, "detectAndCompute")
    k2 = orb.detect(img)
    k2, d2 = orb.compute(img, k2)
    assert_keypoints_equal(k2, ok, "detect->compute")
    assert_descriptors_equal(d2, od, "detect->compute")
    kps, desc, counts = orb.extract_batch([img, img])
    assert_keypoints_equal(kps[1, :counts[1]], ok, "batch")
    assert_descriptors_equal(desc[1, :counts[1]], od, "batch")
    orb.close()


@pytest.mark.parametrize("kind", ["dense", "natural", "flat_then_dense", "mixed"])
def test_fast_density_paths_are_bit_exact(kind):
    """K2 scores a group of pixels either after a compass pre-test (sparse levels) or exhaustively (levels K3 found
    corner-dense in the previous batch, or rows in which nearly every group passes).  Which path runs depends on the
    image content and on the handle's history; the candidate lists must not: every call below -- first call of a fresh
    handle, repeated calls, a switch between frame kinds -- reproduces the oracle's FAST list level by level."""
    w, h = 900, 600
    dense, natural = syn.frame(3, w, h), syn.natural_frame(3, w, h)
    flat = np.full((h, w), 90, np.uint8)
    mixed = natural.copy()
    mixed[:, w // 2:] = dense[:, w // 2:]              # half of every row corner-dense: per-row decisions differ inside a tile
    plan = {"dense": [dense, dense, dense], "natural": [natural, natural], "flat_then_dense": [flat, dense, natural, dense],
            "mixed": [mixed, mixed, dense, mixed]}[kind]
    orb = _orb(1500, HARRIS_SCORE, (w, h))
    P = oracle.Params(nfeatures=1500)
    for step, img in enumerate(plan):
        k, d = orb.detectAndCompute(img)              # runs K2 + K3 on every level: sets the hint for the next call
        ok, od = oracle.detect_and_compute(img, P)
        assert_keypoints_equal(k, ok, "%s step %d" % (kind, step))
        assert_descriptors_equal(d, od, "%s step %d" % (kind, step))
    for level in (0, 3):
        xs, ys, sc = orb.debug_fast_level(plan[-1], level)
        ox, oy, osc = oracle.level_fast(plan[-1], P, level)
        assert np.array_equal(xs, ox) and np.array_equal(ys, oy) and np.array_equal(sc, osc)
    orb.close()
