#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from cv2 (OpenCV) -- the library the reference calls.

The reference (src/FeatureExtractor.cpp:17,19; src/CameraPoseEstimator.cpp:200-213) has no tests or
fixtures for its ORB / BFMatcher path and cannot be built here, so the vectors are
produced by running the same OpenCV calls, in the reference's order, through the cv2 4.13.0 wheel of
the build container:

    orb = cv2.ORB_create(nfeatures=N [, scoreType])      # everything else default == reference defaults
    kps = orb.detect(img, None); kps, desc = orb.compute(img, kps)
    raw = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(q, t, 2);  ratio test in float32

Outputs are stored in canonical keypoint order (octave, y, x) because the order inside a level in OpenCV
is a std::nth_element permutation; descriptors are permuted with their keypoints and the
matcher goldens are computed on the canonically ordered descriptor matrices.

Run (in the build container only; needs cv2):  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from monocular_slam_b200 import synthetic as syn  # noqa: E402

KP = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
               ("class_id", "<i4")])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cv_extract(cv2, img, nf, score_type):
    orb = cv2.ORB_create(nfeatures=nf, scoreType=score_type)
    kp = orb.detect(img, None)
    kp, des = orb.compute(img, kp)
    a = np.zeros(len(kp), KP)
    for i, k in enumerate(kp):
        a[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
    if des is None:
        des = np.zeros((0, 32), np.uint8)
    o = np.lexsort((a["x"], a["y"], a["octave"]))
    return a[o], des[o]


def cv_knn2(cv2, q, t):
    raw = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(q, t, 2)
    idx = -np.ones((len(q), 2), np.int32)
    dist = -np.ones((len(q), 2), np.int32)
    for i, row in enumerate(raw):
        for j, m in enumerate(row):
            assert m.queryIdx == i and m.imgIdx == 0 and m.distance == int(m.distance)
            idx[i, j], dist[i, j] = m.trainIdx, int(m.distance)
    return idx, dist


def cv_ratio(idx, dist, ratio):
    d0 = dist[:, 0].astype(np.float32)
    d1 = dist[:, 1].astype(np.float32)
    keep = (idx[:, 1] >= 0) & (d0 < d1 * np.float32(ratio))
    q = np.nonzero(keep)[0].astype(np.int32)
    return np.stack([q, idx[q, 0], dist[q, 0]], axis=1).astype(np.int32)


def cv_pyramid_hashes(cv2, img, scales):
    h, w = img.shape
    out, prev = [sha(img)], img
    for s in scales[1:]:
        sz = (int(np.rint(np.float32(w) / np.float32(s))), int(np.rint(np.float32(h) / np.float32(s))))
        prev = cv2.resize(prev, sz, interpolation=cv2.INTER_LINEAR_EXACT)
        out.append(sha(prev))
    return out


def main():
    import cv2
    cv2.setNumThreads(1)
    scales = [np.float32(np.float64(np.float32(1.2)) ** l) for l in range(8)]

    # ---- config 1 of BASELINE.json: KITTI-sized frame pair, 2000 kp, both score types, ratio 0.75/0.8/0.85
    big = syn.frame(1, 1241 + 7, 376 + 3)
    f0, f1 = np.ascontiguousarray(big[:376, :1241]), np.ascontiguousarray(big[3:, 7:])
    out = {"canvas": big, "cv2_version": np.array(cv2.__version__)}
    for st, name in ((0, "harris"), (1, "fast")):
        k0, d0 = cv_extract(cv2, f0, 2000, st)
        k1, d1 = cv_extract(cv2, f1, 2000, st)
        idx, dist = cv_knn2(cv2, d0, d1)
        out.update({f"{name}_kp0": k0, f"{name}_desc0": d0, f"{name}_kp1": k1, f"{name}_desc1": d1,
                    f"{name}_knn_idx": idx, f"{name}_knn_dist": dist})
        for r in (0.75, 0.8, 0.85):
            out[f"{name}_good_{int(r * 100)}"] = cv_ratio(idx, dist, r)
    out["pyr_sha_f0"] = np.array(cv_pyramid_hashes(cv2, f0, scales))
    np.savez_compressed(os.path.join(HERE, "kitti_pair.npz"), **out)

    # ---- config 2 shape: 1080p frames from seeds (images are regenerated at test time and checked by sha256)
    out = {}
    for seed in (1, 2):
        img = syn.frame(seed, 1920, 1080)
        k, d = cv_extract(cv2, img, 2000, 0)
        out.update({f"s{seed}_img_sha": np.array(sha(img)), f"s{seed}_kp": k, f"s{seed}_desc": d,
                    f"s{seed}_pyr_sha": np.array(cv_pyramid_hashes(cv2, img, scales))})
    idx, dist = cv_knn2(cv2, out["s1_desc"], out["s2_desc"])
    out.update({"knn_idx": idx, "knn_dist": dist, "good_75": cv_ratio(idx, dist, 0.75)})
    np.savez_compressed(os.path.join(HERE, "hd_frames.npz"), **out)

    # ---- small and degenerate frames (reference default nfeatures = 500)
    out = {}
    cases = {"tiny_97x71": syn.frame(7, 97, 71, nrect=20), "small_200x150": syn.frame(7, 200, 150, nrect=20),
             "odd_333x257": syn.frame(7, 333, 257, nrect=20), "thin_300x63": syn.frame(7, 63, 300, nrect=20),
             "flat_400x300": np.zeros((300, 400), np.uint8),
             "checker_640x480": (((np.indices((480, 640))[0] // 16) + (np.indices((480, 640))[1] // 16)) % 2 * 255).astype(np.uint8),
             "textured_640x480": syn.textured_frame(5, 640, 480)}
    for name, img in cases.items():
        for st, sn in ((0, "harris"), (1, "fast")):
            k, d = cv_extract(cv2, img, 500, st)
            out.update({f"{name}_{sn}_kp": k, f"{name}_{sn}_desc": d})
        out[f"{name}_img"] = img
    # compute() on caller-provided keypoints: shuffled octaves, positions hugging the border (samples leave coarse levels)
    img = syn.frame(3, 800, 600)
    rng = np.random.default_rng(1)
    orb = cv2.ORB_create(nfeatures=700)
    kin = []
    for i in range(400):
        side = int(rng.integers(0, 4))
        x = float(rng.uniform(31, 45)) if side == 0 else float(rng.uniform(755, 768.4)) if side == 1 else float(rng.uniform(31, 768))
        y = float(rng.uniform(31, 45)) if side == 2 else float(rng.uniform(555, 568.4)) if side == 3 else float(rng.uniform(31, 568))
        kin.append(cv2.KeyPoint(x, y, 31.0, float(rng.uniform(0, 360)), 1.0, int(rng.integers(0, 8)), -1))
    kin += [cv2.KeyPoint(10.0, 300.0, 31.0, 45.0, 1.0, 0, -1), cv2.KeyPoint(768.99, 568.99, 37.2, 133.3, 1.0, 3, -1),
            cv2.KeyPoint(31.0, 31.0, 37.2, 33.3, 1.0, 2, -1)]
    kout, dout = orb.compute(img, kin)

    def pack(kps):
        a = np.zeros(len(kps), KP)
        for i, k in enumerate(kps):
            a[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
        return a
    out.update({"border_img": img, "border_kp_in": pack(kin), "border_kp_out": pack(kout), "border_desc": dout})
    np.savez_compressed(os.path.join(HERE, "small_frames.npz"), **out)

    # ---- matcher torture: duplicated rows, all-zero descriptors, Nt in {1,2}, many ties, Nq = 0
    rng = np.random.default_rng(0)
    t = syn.descriptors(1, 3000)
    lo = (rng.integers(0, 2, (500, 32)) * 255).astype(np.uint8)
    mcases = {
        "planted": (syn.planted_queries(2, t, 1000), t),
        "dup_rows": (syn.descriptors(4, 200), np.repeat(syn.descriptors(3, 50), 4, axis=0)),
        "zeros": (np.zeros((10, 32), np.uint8), np.zeros((17, 32), np.uint8)),
        "nt1": (syn.descriptors(5, 7), syn.descriptors(6, 1)),
        "nt2": (syn.descriptors(5, 7), syn.descriptors(6, 2)),
        "low_entropy": (lo[:100], lo),
        "ragged_33x65": (syn.descriptors(8, 33), syn.descriptors(9, 65)),
    }
    out = {}
    for name, (q, tt) in mcases.items():
        idx, dist = cv_knn2(cv2, q, tt)
        out.update({f"{name}_q": q, f"{name}_t": tt, f"{name}_idx": idx, f"{name}_dist": dist})
        if len(tt) >= 2:
            for r in (0.75, 0.8, 0.85):
                out[f"{name}_good_{int(r * 100)}"] = cv_ratio(idx, dist, r)
    np.savez_compressed(os.path.join(HERE, "matcher_cases.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
