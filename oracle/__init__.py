"""ctypes wrapper around oracle/liborb_oracle.so (CPU restatement of the reference's ORB + BF-Hamming path and of the
fundamental-matrix outlier filter that follows it: orb_oracle.c, fmat_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of oracle/orb_oracle.c.  The product package
(monocular_slam_b200) never imports this module; tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do, as the checker or reported baseline.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborb_oracle.so")

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")]
)
assert KEYPOINT_DTYPE.itemsize == 28

HARRIS_SCORE, FAST_SCORE = 0, 1


class Params(C.Structure):
    """cv::ORB constructor arguments; defaults are the reference's (src/FeatureExtractor.h:23-24)."""

    _fields_ = [
        ("nfeatures", C.c_int32),
        ("scale_factor", C.c_float),
        ("nlevels", C.c_int32),
        ("edge_threshold", C.c_int32),
        ("first_level", C.c_int32),
        ("wta_k", C.c_int32),
        ("score_type", C.c_int32),
        ("patch_size", C.c_int32),
        ("fast_threshold", C.c_int32),
    ]

    def __init__(self, nfeatures=500, scale_factor=1.2, nlevels=8, edge_threshold=31, first_level=0, wta_k=2,
                 score_type=HARRIS_SCORE, patch_size=31, fast_threshold=20):
        super().__init__(nfeatures, scale_factor, nlevels, edge_threshold, first_level, wta_k, score_type, patch_size,
                         fast_threshold)


def build(force=False):
    """Compile liborb_oracle.so with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("orb_oracle.c", "fmat_oracle.c", "tri_oracle.c", "loop_oracle.c", "bow_oracle.c", "jpeg_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, i32p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_float)
        PP = C.POINTER(Params)
        L.orc_level_scale.restype = C.c_float
        L.orc_level_scale.argtypes = [PP, C.c_int]
        L.orc_level_size.argtypes = [PP, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_level_quotas.argtypes = [PP, C.POINTER(C.c_int)]
        L.orc_resize_linear_exact.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int]
        L.orc_fast_score_map.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_fast_nms.argtypes = [u8p, C.c_int, C.c_int, C.c_int, i32p, i32p, i32p, C.c_int]
        L.orc_harris.restype = C.c_float
        L.orc_harris.argtypes = [u8p, C.c_int, C.c_int, C.c_int]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_ic_angle.restype = C.c_float
        L.orc_ic_angle.argtypes = [u8p, C.c_int, C.c_int, C.c_int]
        L.orc_gauss_kernel7.argtypes = [f32p]
        L.orc_blur7.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_describe.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, u8p]
        L.orc_pyramid_level.argtypes = [PP, u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, u8p]
        L.orc_detect.argtypes = [PP, u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int]
        L.orc_compute.argtypes = [PP, u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, u8p]
        L.orc_level_fast.argtypes = [PP, u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, i32p, i32p, i32p, C.c_int]
        L.orc_knn2.argtypes = [u8p, C.c_int64, u8p, C.c_int64, i32p, i32p]
        L.orc_bgr2gray.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p]
        L.orc_bgr2gray.restype = None
        L.orc_ratio_test.restype = C.c_int64
        L.orc_ratio_test.argtypes = [i32p, i32p, C.c_int64, C.c_float, i32p, i32p, i32p]
        f64p = C.POINTER(C.c_double)
        L.orc_solve_cubic.argtypes = [f64p, f64p]
        L.orc_fm_7point.argtypes = [f32p, f32p, f64p]
        L.orc_fm_8point.argtypes = [f32p, f32p, C.c_int, f64p]
        L.orc_fm_errors.argtypes = [f32p, f32p, C.c_int, f64p, f32p]
        L.orc_fm_errors.restype = None
        L.orc_fm_ransac.argtypes = [f32p, f32p, C.c_int, C.c_double, C.c_double, C.c_int, u8p, f64p, i32p]
        L.orc_compute_fundamental.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                              u8p, f64p]
        L.orc_triangulate.argtypes = [f64p, f64p, C.c_int, f64p, f64p, f64p, f64p, f64p, u8p]
        L.orc_triangulate_best.argtypes = [f64p, f64p, C.c_int, f64p, f64p, C.c_int, f64p, f64p, f64p, i32p]
        L.orc_associate.argtypes = [C.c_void_p, i32p, C.c_int, C.c_int, i32p, C.c_int, i32p, i32p, i32p]
        L.orc_select_new.argtypes = [C.c_void_p, i32p, C.c_int, C.c_int, i32p, i32p, C.c_int32, u8p]
        L.orc_nbest.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, i32p, i32p]
        L.orc_nbest.restype = None
        L.orc_loop_score.argtypes = [u8p, C.c_int, u8p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
        u32p = C.POINTER(C.c_uint32)
        L.bow_voc_create.restype = C.c_void_p
        L.bow_voc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, u8p, u8p, f64p]
        L.bow_voc_free.argtypes = [C.c_void_p]
        L.bow_voc_free.restype = None
        L.bow_voc_words.argtypes = [C.c_void_p]
        L.bow_transform_feature.argtypes = [C.c_void_p, u8p, C.c_int, u32p, f64p, u32p]
        L.bow_transform_feature.restype = None
        L.bow_transform.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, u32p, f64p, C.POINTER(C.c_int), u32p, i32p, u32p, C.POINTER(C.c_int)]
        L.bow_transform.restype = None
        L.bow_score.restype = C.c_double
        L.bow_score.argtypes = [C.c_int, u32p, f64p, C.c_int, u32p, f64p, C.c_int]
        L.bow_stop_words.argtypes = [C.c_void_p, C.c_double]
        L.bow_parent_node.restype = C.c_uint32
        L.bow_parent_node.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        L.bow_word_weight.restype = C.c_double
        L.bow_word_weight.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_jpeg_probe.argtypes = [u8p, C.c_size_t, i32p]
        L.orc_jpeg_decode_gray.argtypes = [u8p, C.c_size_t, u8p, C.c_int]
        L.orc_jpeg_probe_colour.argtypes = [u8p, C.c_size_t, i32p]
        L.orc_jpeg_decode_bgr.argtypes = [u8p, C.c_size_t, u8p, C.c_int]
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def bgr2gray(img):
    """cvtColor(BGR2GRAY) as cv::ORB applies it to 3-channel input (orc_bgr2gray)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 3 and img.shape[2] == 3
    out = np.zeros(img.shape[:2], np.uint8)
    lib().orc_bgr2gray(_u8(img), img.shape[1], img.shape[0], img.strides[0], _u8(out))
    return out


def _gray(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim == 3:
        return bgr2gray(img)
    assert img.ndim == 2
    return img


# ---------------------------------------------------------------- geometry
def level_scales(params):
    return np.array([lib().orc_level_scale(C.byref(params), l) for l in range(params.nlevels)], dtype=np.float32)


def level_sizes(params, w, h):
    out = []
    for l in range(params.nlevels):
        lw, lh = C.c_int(), C.c_int()
        lib().orc_level_size(C.byref(params), w, h, l, C.byref(lw), C.byref(lh))
        out.append((lw.value, lh.value))
    return out


def level_quotas(params):
    q = (C.c_int * params.nlevels)()
    lib().orc_level_quotas(C.byref(params), q)
    return list(q)


# ---------------------------------------------------------------- stages
def resize_linear_exact(src, dw, dh):
    src = _gray(src)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_exact(_u8(src), src.shape[1], src.shape[0], src.shape[1], _u8(dst), dw, dh, dw)
    return dst


def pyramid_level(params, img, level):
    img = _gray(img)
    h, w = img.shape
    lw, lh = level_sizes(params, w, h)[level]
    out = np.empty((lh, lw), np.uint8)
    rc = lib().orc_pyramid_level(C.byref(params), _u8(img), w, h, w, level, _u8(out))
    assert rc == 0
    return out


def fast_score_map(img, threshold=20):
    img = _gray(img)
    h, w = img.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_fast_score_map(_u8(img), w, h, w, threshold, _u8(out))
    return out


def fast_nms(score, border=31):
    score = _gray(score)
    h, w = score.shape
    cap = (w // 2 + 1) * (h // 2 + 1)
    xs, ys, sc = (np.empty(cap, np.int32) for _ in range(3))
    n = lib().orc_fast_nms(_u8(score), w, h, border, _i32(xs), _i32(ys), _i32(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def harris(img, x, y):
    img = _gray(img)
    return np.float32(lib().orc_harris(_u8(img), img.shape[1], int(x), int(y)))


def fast_atan2(y, x):
    return np.float32(lib().orc_fast_atan2(float(np.float32(y)), float(np.float32(x))))


def ic_angle(img, x, y):
    img = _gray(img)
    return np.float32(lib().orc_ic_angle(_u8(img), img.shape[1], int(x), int(y)))


def gauss_kernel7():
    k = np.empty(7, np.float32)
    lib().orc_gauss_kernel7(k.ctypes.data_as(C.POINTER(C.c_float)))
    return k


def blur7(img):
    img = _gray(img)
    out = np.empty_like(img)
    lib().orc_blur7(_u8(img), img.shape[1], img.shape[0], img.shape[1], _u8(out))
    return out


# ---------------------------------------------------------------- whole path
def detect(img, params, cap=None):
    """ORB detect as the reference calls it (src/FeatureExtractor.cpp:17); canonical order (octave, y, x)."""
    img = _gray(img)
    h, w = img.shape
    cap = cap or max(4 * params.nfeatures + 4096, 8192)
    kps = np.zeros(cap, KEYPOINT_DTYPE)
    n = lib().orc_detect(C.byref(params), _u8(img), w, h, w, kps.ctypes.data, cap)
    if n < 0:
        raise ValueError("oracle: unsupported ORB parameters (%d)" % n)
    if n > cap:
        return detect(img, params, cap=n)
    return kps[:n].copy()


def compute(img, kps, params):
    """ORB compute (src/FeatureExtractor.cpp:19): returns (filtered/regrouped keypoints, N x 32 descriptors)."""
    img = _gray(img)
    h, w = img.shape
    kps = np.array(kps, dtype=KEYPOINT_DTYPE, copy=True)
    desc = np.zeros((len(kps), 32), np.uint8)
    m = lib().orc_compute(C.byref(params), _u8(img), w, h, w, kps.ctypes.data, len(kps), _u8(desc))
    if m < 0:
        raise ValueError("oracle: bad keypoints / parameters (%d)" % m)
    return kps[:m].copy(), desc[:m].copy()


def detect_and_compute(img, params):
    return compute(img, detect(img, params), params)


def level_fast(img, params, level):
    img = _gray(img)
    h, w = img.shape
    lw, lh = level_sizes(params, w, h)[level]
    cap = (lw // 2 + 1) * (lh // 2 + 1)
    xs, ys, sc = (np.empty(cap, np.int32) for _ in range(3))
    n = lib().orc_level_fast(C.byref(params), _u8(img), w, h, w, level, _i32(xs), _i32(ys), _i32(sc), cap)
    assert n >= 0
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def knn2(q, t):
    """BFMatcher(NORM_HAMMING,false).knnMatch(q,t,2) (src/CameraPoseEstimator.cpp:202-204): (idx, dist) each nq x 2, -1 = absent."""
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx = np.empty((len(q), 2), np.int32)
    dist = np.empty((len(q), 2), np.int32)
    lib().orc_knn2(_u8(q), len(q), _u8(t), len(t), _i32(idx), _i32(dist))
    return idx, dist


def ratio_test(idx, dist, ratio):
    """Lowe ratio exactly as src/CameraPoseEstimator.cpp:208-212: returns (query_idx, train_idx, distance) arrays."""
    idx = np.ascontiguousarray(idx, np.int32)
    dist = np.ascontiguousarray(dist, np.int32)
    nq = len(idx)
    gq, gt, gd = (np.empty(nq, np.int32) for _ in range(3))
    n = lib().orc_ratio_test(_i32(idx), _i32(dist), nq, float(ratio), _i32(gq), _i32(gt), _i32(gd))
    return gq[:n].copy(), gt[:n].copy(), gd[:n].copy()


def match_features(d1, d2, ratio=0.8):
    """matchFeatures(descriptors1, descriptors2, matches, ratio) of src/CameraPoseEstimator.cpp:200-213."""
    idx, dist = knn2(d1, d2)
    return ratio_test(idx, dist, ratio)


# ---- fundamental-matrix outlier filter (fmat_oracle.c) ----

def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f64(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pts(p):
    p = np.ascontiguousarray(p, np.float32).reshape(-1, 2)
    return p


def solve_cubic(coeffs):
    """cv::solveCubic: real roots in OpenCV's order."""
    c = np.ascontiguousarray(coeffs, np.float64)
    x = np.zeros(3)
    n = lib().orc_solve_cubic(_f64(c), _f64(x))
    return x[:max(n, 0)].copy()


def fm_7point(m1, m2):
    """run7Point on 7 correspondences: (k, 3, 3) candidate matrices, k <= 3."""
    m1, m2 = _pts(m1), _pts(m2)
    assert len(m1) == 7 and len(m2) == 7
    out = np.zeros(27)
    n = lib().orc_fm_7point(_f32(m1), _f32(m2), _f64(out))
    return out[:9 * n].reshape(n, 3, 3).copy()


def fm_8point(m1, m2):
    """findFundamentalMat(m1, m2, FM_8POINT) (src/CameraPoseEstimator.cpp:585): 3x3 or None."""
    m1, m2 = _pts(m1), _pts(m2)
    F = np.zeros(9)
    ok = lib().orc_fm_8point(_f32(m1), _f32(m2), len(m1), _f64(F))
    return F.reshape(3, 3) if ok else None


def fm_errors(m1, m2, F):
    m1, m2 = _pts(m1), _pts(m2)
    F = np.ascontiguousarray(F, np.float64).reshape(9)
    err = np.empty(len(m1), np.float32)
    lib().orc_fm_errors(_f32(m1), _f32(m2), len(m1), _f64(F), _f32(err))
    return err


def fm_ransac(m1, m2, thr=3.0, conf=0.85, max_iters=1000):
    """findFundamentalMat(m1, m2, FM_RANSAC, thr, conf, status) (src/CameraPoseEstimator.cpp:563) for >= 8 points:
    (F or None, mask uint8[n], iterations run)."""
    m1, m2 = _pts(m1), _pts(m2)
    n = len(m1)
    assert n >= 8
    mask = np.zeros(n, np.uint8)
    F = np.zeros(9)
    info = np.zeros(2, np.int32)
    ok = lib().orc_fm_ransac(_f32(m1), _f32(m2), n, float(thr), float(conf), int(max_iters), _u8(mask), _f64(F), _i32(info))
    return (F.reshape(3, 3) if ok else None), mask, int(info[0])


def compute_fundamental(pos1, pos2, matches, thr=3.0, conf=0.85):
    """computeFundamentalMatrix (src/CameraPoseEstimator.cpp:545-586): pos1/pos2 float32 (n, 2) keypoint positions,
    matches a structured/int array of (query_idx, train_idx, img_idx, distance) rows (16 bytes each).
    Returns (F 3x3 (zeros if none), status uint8[len(matches)], ninliers)."""
    pos1, pos2 = _pts(pos1), _pts(pos2)
    matches = np.ascontiguousarray(matches)
    assert matches.dtype.itemsize * (matches.shape[1] if matches.ndim == 2 else 1) == 16
    nm = len(matches)
    status = np.zeros(max(nm, 1), np.uint8)
    F = np.zeros(9)
    ninl = lib().orc_compute_fundamental(_f32(pos1), 2, _f32(pos2), 2, matches.ctypes.data, nm, float(thr), float(conf),
                                         _u8(status), _f64(F))
    return F.reshape(3, 3), status[:nm], int(ninl)


# ------------------------------------------------------------------------------------------------ triangulation / association
def _d(a, shape=None):
    a = np.ascontiguousarray(a, np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def triangulate(pts1, pts2, Rt1, Rt2, K1, K2):
    """TriangulateMultiplePointsFromTwoView (src/CameraPoseEstimator.cpp:134-152) with countFront: (X[n,3], front[n] uint8, count)."""
    p1, pp1 = _d(pts1)
    p2, pp2 = _d(pts2)
    n = len(p1)
    X = np.zeros((n, 3), np.float64)
    front = np.zeros(max(n, 1), np.uint8)
    cnt = lib().orc_triangulate(pp1, pp2, n, _d(Rt1, 12)[1], _d(Rt2, 12)[1], _d(K1, 9)[1], _d(K2, 9)[1],
                                X.ctypes.data_as(C.POINTER(C.c_double)), _u8(front))
    return X, front[:n], int(cnt)


def triangulate_best(pts1, pts2, Rt1, Rts, K1, K2):
    """The four-hypothesis test of the bootstrap (src/CameraPoseEstimator.cpp:334-349): (best index, counts[nhyp], X of the winner)."""
    p1, pp1 = _d(pts1)
    p2, pp2 = _d(pts2)
    R, pR = _d(Rts)
    nh = R.size // 12
    n = len(p1)
    X = np.zeros((n, 3), np.float64)
    counts = np.zeros(nh, np.int32)
    best = lib().orc_triangulate_best(pp1, pp2, n, _d(Rt1, 12)[1], pR, nh, _d(K1, 9)[1], _d(K2, 9)[1],
                                      X.ctypes.data_as(C.POINTER(C.c_double)), _i32(counts))
    return int(best), counts, X


_DMATCH = np.dtype([("query_idx", "<i4"), ("train_idx", "<i4"), ("img_idx", "<i4"), ("distance", "<f4")])


def associate(matches, counts, premap, ncur):
    """The association loop (src/CameraPoseEstimator.cpp:402-455) of one frame: matches [back, cap] DMatch (inliers only),
    counts [back], premap [back, cap] int32.  Returns (cur_map[ncur], assoc_q[k], assoc_mp[k])."""
    m = np.ascontiguousarray(matches, _DMATCH)
    back, cap = m.shape
    n = np.ascontiguousarray(counts, np.int32)
    pm = np.ascontiguousarray(premap, np.int32)
    assert pm.shape == (back, cap)
    cur = np.zeros(max(ncur, 1), np.int32)
    aq, amp = np.zeros(max(ncur, 1), np.int32), np.zeros(max(ncur, 1), np.int32)
    k = lib().orc_associate(m.ctypes.data, _i32(n), back, cap, _i32(pm), ncur, _i32(cur), _i32(aq), _i32(amp))
    return cur[:ncur], aq[:k], amp[:k]


def select_new(matches, counts, premap, cur_map, next_id=0):
    """The new-map-point loop (src/CameraPoseEstimator.cpp:488-512): returns (accept[back, cap] uint8, premap', cur_map', n_new)."""
    m = np.ascontiguousarray(matches, _DMATCH)
    back, cap = m.shape
    n = np.ascontiguousarray(counts, np.int32)
    pm = np.array(premap, np.int32, copy=True)
    cm = np.array(cur_map, np.int32, copy=True)
    acc = np.zeros((back, cap), np.uint8)
    k = lib().orc_select_new(m.ctypes.data, _i32(n), back, cap, _i32(pm), _i32(cm), int(next_id), _u8(acc))
    return acc, pm, cm, int(k)


# ------------------------------------------------------------------------------------------------ loop-closure scoring
def nbest(d1, d2, n=10):
    """LoopCloser::NBestMatches (src/LoopCloser.cpp:53-105) with the Hamming distance: (dist[nq, n] int32, idx[nq, n] int32),
    ascending, absent entries (fewer than n train rows) dist = -1 / idx = -1."""
    q, t = np.ascontiguousarray(d1, np.uint8), np.ascontiguousarray(d2, np.uint8)
    dist = np.zeros((len(q), n), np.int32)
    idx = np.zeros((len(q), n), np.int32)
    if len(q):
        lib().orc_nbest(_u8(q), len(q), _u8(t), len(t), n, _i32(dist), _i32(idx))
    return dist, idx


def loop_score(d1, frames_desc, frame_counts, n=10, thr=40):
    """LoopCloser::DetectLoop's scoring (src/LoopCloser.cpp:19-51): for every stored frame the number of n-best distances
    below thr; returns (counts[nframes] int32, index of the first maximum or -1)."""
    q = np.ascontiguousarray(d1, np.uint8)
    fd = np.ascontiguousarray(frames_desc, np.uint8)
    nf, cap = fd.shape[0], fd.shape[1]
    fc = np.ascontiguousarray(frame_counts, np.int32)
    counts = np.zeros(max(nf, 1), np.int32)
    best = lib().orc_loop_score(_u8(q), len(q), _u8(fd), _i32(fc), nf, cap, n, thr, _i32(counts))
    return counts[:nf], int(best)


# ---- bag of words (bow_oracle.c; ThirdParty/DBoW2)
TF_IDF, TF, IDF, BINARY = 0, 1, 2, 3
L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = 0, 1, 2, 3, 4, 5


def _u32(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


class BowVocabulary:
    """TemplatedVocabulary<FORB::TDescriptor, FORB> over the arrays of a text vocabulary (synthetic.vocabulary /
    read_vocabulary_text): transform, score, stopWords, getParentNode, getWordWeight as bow_oracle.c restates them."""

    def __init__(self, voc, scoring=L1_NORM, weighting=TF_IDF):
        self.scoring, self.weighting = int(voc.get("scoring", scoring)), int(voc.get("weighting", weighting))
        parent = np.ascontiguousarray(voc["parent"], np.int32)
        leaf = np.ascontiguousarray(voc["leaf"], np.uint8)
        desc = np.ascontiguousarray(voc["desc"], np.uint8)
        weight = np.ascontiguousarray(voc["weight"], np.float64)
        self._h = lib().bow_voc_create(int(voc["k"]), int(voc["L"]), self.scoring, self.weighting, len(parent), _i32(parent), _u8(leaf),
                                       _u8(desc), _f64(weight))
        if not self._h:
            raise ValueError("bow_voc_create: parent[i] must be in [0, i)")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().bow_voc_free(self._h)
            self._h = None

    def size(self):
        return lib().bow_voc_words(self._h)

    def transform_features(self, desc, levelsup=0):
        """Per feature: (word id, word weight, node id `levelsup` levels above the words)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word, weight, nid = np.zeros(n, np.uint32), np.zeros(n, np.float64), np.zeros(n, np.uint32)
        for i in range(n):
            lib().bow_transform_feature(self._h, _u8(desc[i]), levelsup, _u32(word[i:]), _f64(weight[i:]), _u32(nid[i:]))
        return word, weight, nid

    def transform(self, desc, levelsup=None):
        """(words, values) of the BowVector; with ``levelsup`` also (nodes, offsets, features) of the FeatureVector."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        m = max(n, 1)
        words, vals, nb = np.zeros(m, np.uint32), np.zeros(m, np.float64), C.c_int(0)
        if levelsup is None:
            lib().bow_transform(self._h, _u8(desc), n, 0, _u32(words), _f64(vals), C.byref(nb), None, None, None, None)
            return words[:nb.value].copy(), vals[:nb.value].copy()
        nodes, offs, feats, nf = np.zeros(m, np.uint32), np.zeros(m + 1, np.int32), np.zeros(m, np.uint32), C.c_int(0)
        lib().bow_transform(self._h, _u8(desc), n, int(levelsup), _u32(words), _f64(vals), C.byref(nb), _u32(nodes), _i32(offs), _u32(feats),
                            C.byref(nf))
        k = nf.value
        return words[:nb.value].copy(), vals[:nb.value].copy(), nodes[:k].copy(), offs[:k + 1].copy(), feats[:offs[k]].copy()

    def score(self, a, b):
        return bow_score(self.scoring, a, b)

    def stop_words(self, min_weight):
        return lib().bow_stop_words(self._h, float(min_weight))

    def parent_node(self, wid, levelsup):
        return int(lib().bow_parent_node(self._h, int(wid), int(levelsup)))

    def word_weight(self, wid):
        return float(lib().bow_word_weight(self._h, int(wid)))


def bow_score(scoring, a, b):
    """score(v1, v2) of the scoring object `scoring`; a, b = (words ascending, values)."""
    w1, v1 = np.ascontiguousarray(a[0], np.uint32), np.ascontiguousarray(a[1], np.float64)
    w2, v2 = np.ascontiguousarray(b[0], np.uint32), np.ascontiguousarray(b[1], np.float64)
    return float(lib().bow_score(int(scoring), _u32(w1), _f64(v1), len(w1), _u32(w2), _f64(v2), len(w2)))


# ---- grey-scale baseline JPEG (jpeg_oracle.c; the decoder behind imread in src/FrameLoader.cpp:62)
JPEG_UNSUPPORTED, JPEG_CORRUPT = -1, -2


def jpeg_probe(data):
    """(width, height, restart interval, blocks) of a file the decoder handles; raises ValueError with the status otherwise."""
    buf = np.frombuffer(bytes(data), np.uint8)
    info = np.zeros(4, np.int32)
    rc = lib().orc_jpeg_probe(_u8(buf), len(buf), _i32(info))
    if rc:
        raise ValueError(rc)
    return tuple(int(x) for x in info)


def jpeg_decode_gray(data):
    w, h, _, _ = jpeg_probe(data)
    buf = np.frombuffer(bytes(data), np.uint8)
    out = np.zeros((h, w), np.uint8)
    rc = lib().orc_jpeg_decode_gray(_u8(buf), len(buf), _u8(out), w)
    if rc:
        raise ValueError(rc)
    return out


def jpeg_decode_bgr(data):
    """A three-component (YCbCr 4:2:0 / 4:2:2 / 4:4:4) baseline file as imread returns it: [h][w][3] BGR."""
    buf = np.frombuffer(bytes(data), np.uint8)
    info = np.zeros(12, np.int32)
    rc = lib().orc_jpeg_probe_colour(_u8(buf), len(buf), _i32(info))
    if rc:
        raise ValueError(rc)
    w, h = int(info[0]), int(info[1])
    out = np.zeros((h, w, 3), np.uint8)
    rc = lib().orc_jpeg_decode_bgr(_u8(buf), len(buf), _u8(out), 3 * w)
    if rc:
        raise ValueError(rc)
    return out
