"""GPU: the ingest ring (encoded frames -> host decode workers -> pinned slots -> pipelined extract + match) must give, frame by
frame, what the oracle gives on the decoded pixels -- the reference's FrameLoader (imread, src/FrameLoader.cpp:62) followed by
FeatureExtractor::process and matchFeatures.  PNG is lossless, so the decoded pixels are the original ones."""
import numpy as np
import pytest

import oracle
from helpers import assert_descriptors_equal, assert_keypoints_equal
from monocular_slam_b200 import ORB, BFMatcher
from monocular_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("channels", [1, 3])
def test_ring_matches_oracle(channels, tmp_path):
    from monocular_slam_b200.ingest import IngestRing
    w, h, n = 640, 480, 11
    if channels == 1:
        frames = list(syn.sequence(n, w, h, seed=77))
    else:
        frames = [syn.bgr_frame(900 + i, w, h) for i in range(n)]
    sources = []
    for i, f in enumerate(frames):          # half of them as files, half as encoded bytes
        if i % 2:
            p = tmp_path / ("f%03d.png" % i)
            assert cv2.imwrite(str(p), f)
            sources.append(str(p))
        else:
            sources.append(cv2.imencode(".png", f)[1].tobytes())
    orb = ORB(nfeatures=600, max_size=(w, h), max_batch=4)
    bf = BFMatcher()
    ring = IngestRing(orb, bf, w, h, batch=4, workers=3, ratio=0.8, channels=channels)
    P = oracle.Params(nfeatures=600)
    prev = None
    seen = 0
    for first, nb, kps, desc, counts, good, ngood in ring.run(sources):
        assert first == seen
        for i in range(nb):
            gray = frames[first + i] if channels == 1 else oracle.bgr2gray(frames[first + i])
            ok, od = oracle.detect_and_compute(gray, P)
            c = int(counts[i])
            assert_keypoints_equal(kps[i, :c], ok, "frame %d" % (first + i))
            assert_descriptors_equal(desc[i, :c], od, "frame %d" % (first + i))
            if prev is not None:
                gq, gt, gd = oracle.match_features(od, prev, 0.8)
                g = good[i, :int(ngood[i])]
                assert np.array_equal(g["query_idx"], gq) and np.array_equal(g["train_idx"], gt)
            prev = od
        seen += nb
    assert seen == n and ring.stats["frames"] == n and ring.stats["decode_seconds"] > 0
    with pytest.raises(ValueError):
        list(ring.run([cv2.imencode(".png", np.zeros((10, 10), np.uint8))[1].tobytes()]))
    ring.close()
    bf.close()
    orb.close()
