"""Shared helpers for the parity tests (the oracle is the checker; the product is monocular_slam_b200)."""
import hashlib

import numpy as np


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def assert_keypoints_equal(got, want, what=""):
    assert len(got) == len(want), "%s: %d keypoints, expected %d" % (what, len(got), len(want))
    for f in ("octave", "x", "y", "size", "response", "angle", "class_id"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, "%s: field %s differs at %d rows (first %d: got %r want %r)" % (
            what, f, bad.size, bad[0], got[f][bad[0]], want[f][bad[0]])


def assert_descriptors_equal(got, want, what=""):
    assert got.shape == want.shape, "%s: descriptor shape %s, expected %s" % (what, got.shape, want.shape)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, "%s: %d of %d descriptors differ (first row %d)" % (what, bad.size, len(want), bad[0] if bad.size else -1)


def knn_arrays(matches, counts):
    """(matches[nq,2] DMatch, counts) -> idx[nq,2], dist[nq,2] with -1 for absent entries (oracle layout)."""
    idx = matches["train_idx"].astype(np.int32).copy()
    dist = matches["distance"].astype(np.int32).copy()
    assert np.array_equal(matches["distance"], dist.astype(np.float32)), "distances must be integral floats"
    for j in range(2):
        absent = counts <= j
        idx[absent, j] = -1
        dist[absent, j] = -1
    return idx, dist
