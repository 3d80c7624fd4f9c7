"""GPU parity: sequence mode (batched extraction + consecutive-frame matching with descriptors left on the device)
against the oracle run frame by frame -- the reference's FeatureExtractor::process followed by
matchFeatures(desc_cur, desc_prev) (src/FeatureExtractor.cpp:13-31, src/CameraPoseEstimator.cpp:200-213,409)."""
import numpy as np
import pytest

import oracle
from helpers import assert_descriptors_equal, assert_keypoints_equal
from monocular_slam_b200 import ORB, BFMatcher, DMATCH_DTYPE, KEYPOINT_DTYPE
from monocular_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _oracle_sequence(seq, nf, ratio):
    P = oracle.Params(nfeatures=nf)
    ext = [oracle.detect_and_compute(f, P) for f in seq]
    matches = [None] + [oracle.match_features(ext[i][1], ext[i - 1][1], ratio) for i in range(1, len(seq))]
    return ext, matches


def _check_matches(good_row, n, want):
    q, t, d = want
    assert n == len(q)
    g = good_row[:n]
    assert np.array_equal(g["query_idx"], q) and np.array_equal(g["train_idx"], t)
    assert np.array_equal(g["distance"].astype(np.int32), d) and (g["img_idx"] == 0).all()


@pytest.mark.parametrize("kernel", ["auto", "integer", "tensor"])
def test_sequence_host_path(kernel):
    """Batched frame-pair matching runs on the tensor-core kernel when the pairs are large enough, on the integer-pipe kernel
    otherwise (hamx_set_kernel): all three settings must reproduce the oracle."""
    from monocular_slam_b200 import _lib
    seq = syn.sequence(7, 800, 600, seed=21)
    ext, matches = _oracle_sequence(seq, 1000, 0.8)
    orb = ORB(nfeatures=1000, max_size=(800, 600), max_batch=4)
    m = BFMatcher()
    m.set_kernel({"auto": _lib.KERNEL_AUTO, "integer": _lib.KERNEL_INTEGER, "tensor": _lib.KERNEL_TENSOR}[kernel])
    cap = orb.default_cap
    # two submissions (4 + 3 frames): frame 4 must be matched against frame 3 of the previous batch
    done = 0
    for chunk in (list(seq[:4]), list(seq[4:])):
        kps, desc, counts = orb.extract_batch(chunk, cap=cap)
        good, ngood = orb.match_consecutive(m, 0.8, cap, len(chunk))
        for i in range(len(chunk)):
            f = done + i
            assert_keypoints_equal(kps[i, :counts[i]], ext[f][0], "frame %d" % f)
            assert_descriptors_equal(desc[i, :counts[i]], ext[f][1], "frame %d" % f)
            if f == 0:
                assert ngood[i] == 0
            else:
                _check_matches(good[i], int(ngood[i]), matches[f])
        done += len(chunk)
    orb.reset_sequence()
    kps, desc, counts = orb.extract_batch(list(seq[:2]), cap=cap)
    good, ngood = orb.match_consecutive(m, 0.8, cap, 2)
    assert ngood[0] == 0
    _check_matches(good[1], int(ngood[1]), matches[1])
    m.close()
    orb.close()


@pytest.mark.parametrize("w,h,nf", [(1241, 376, 2000), (640, 480, 700)])   # byte-wise and 16-byte-vector ingest
def test_sequence_device_path(w, h, nf):
    import torch
    seq = syn.sequence(5, w, h, seed=22)
    ext, matches = _oracle_sequence(seq, nf, 0.75)
    B, (H, W) = len(seq), seq[0].shape
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        orb = ORB(nfeatures=nf, max_size=(W, H), max_batch=B)
        m = BFMatcher()
        orb.set_stream(stream.cuda_stream)
        m.set_stream(stream.cuda_stream)
        cap = orb.default_cap
        d_frames = torch.from_numpy(seq).cuda()
        d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device="cuda")
        d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
        d_good = torch.empty((B, cap, 4), dtype=torch.int32, device="cuda")
        d_ngood = torch.zeros(B, dtype=torch.int64, device="cuda")
        for _ in range(2):   # twice: counters, arrival counters and workspaces must be reusable
            orb.extract_batch_dev(d_frames.data_ptr(), W * H, B, W, H, W, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_cnt.data_ptr())
            m.match_consecutive_dev(d_desc.data_ptr(), d_cnt.data_ptr(), B, cap, 0, 0, 0.75, d_good.data_ptr(), d_ngood.data_ptr())
        orb.check_dev()
        stream.synchronize()
    cnt = d_cnt.cpu().numpy()
    kps = d_kps.cpu().numpy().view(KEYPOINT_DTYPE).reshape(B, cap)
    desc = d_desc.cpu().numpy()
    good = d_good.cpu().numpy().view(DMATCH_DTYPE).reshape(B, cap)
    ngood = d_ngood.cpu().numpy()
    for f in range(B):
        assert_keypoints_equal(kps[f, :cnt[f]], ext[f][0], "frame %d" % f)
        assert_descriptors_equal(desc[f, :cnt[f]], ext[f][1], "frame %d" % f)
        if f:
            _check_matches(good[f], int(ngood[f]), matches[f])
    assert ngood[0] == 0
    m.close()
    orb.close()


def test_batched_pairs_ragged():
    """hamx_match_pairs_dev on pairs of very different sizes, including empty ones."""
    import torch
    sizes = [(300, 500), (1, 1), (0, 40), (40, 0), (257, 129), (1000, 3), (5, 2000)]
    sets = [(syn.descriptors(2 * i + 1, nq), syn.descriptors(2 * i + 2, nt)) for i, (nq, nt) in enumerate(sizes)]
    max_nq = max(s[0] for s in sizes)
    max_nt = max(s[1] for s in sizes)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        m = BFMatcher()
        m.set_stream(stream.cuda_stream)
        dq = [torch.from_numpy(q).cuda() if len(q) else torch.zeros((1, 32), dtype=torch.uint8, device="cuda") for q, _ in sets]
        dt = [torch.from_numpy(t).cuda() if len(t) else torch.zeros((1, 32), dtype=torch.uint8, device="cuda") for _, t in sets]
        table = np.zeros(len(sets), np.dtype([("q", "<u8"), ("t", "<u8"), ("nq", "<i4"), ("nt", "<i4")]))
        for i, (q, t) in enumerate(sets):
            table[i] = (dq[i].data_ptr(), dt[i].data_ptr(), len(q), len(t))
        d_pairs = torch.from_numpy(table.view(np.uint8)).cuda()
        d_good = torch.empty((len(sets), max_nq, 4), dtype=torch.int32, device="cuda")
        d_ngood = torch.zeros(len(sets), dtype=torch.int64, device="cuda")
        m.match_pairs_dev(d_pairs.data_ptr(), len(sets), max_nq, max_nt, 0.8, d_good.data_ptr(), max_nq, d_ngood.data_ptr())
        stream.synchronize()
    good = d_good.cpu().numpy().view(DMATCH_DTYPE).reshape(len(sets), max_nq)
    ngood = d_ngood.cpu().numpy()
    for i, (q, t) in enumerate(sets):
        _check_matches(good[i], int(ngood[i]), oracle.match_features(q, t, 0.8))
    m.close()


def test_pipelined_submit_wait():
    """orbx_submit_batch / orbx_wait_batch: two batches in flight, results identical to the frame-by-frame oracle,
    including the match of each batch's first frame against the previous batch's last frame."""
    import torch
    seq = syn.sequence(8, 800, 600, seed=23)
    ext, matches = _oracle_sequence(seq, 1000, 0.8)
    orb = ORB(nfeatures=1000, max_size=(800, 600), max_batch=3)
    m = BFMatcher()
    cap = orb.default_cap

    def buffers(n):
        return (torch.zeros((n, cap, 7), dtype=torch.float32).pin_memory().numpy().view(KEYPOINT_DTYPE).reshape(n, cap),
                torch.zeros((n, cap, 32), dtype=torch.uint8).pin_memory().numpy(),
                np.zeros(n, np.int32),
                torch.zeros((n, cap, 4), dtype=torch.int32).pin_memory().numpy().view(DMATCH_DTYPE).reshape(n, cap),
                np.zeros(n, np.int64))

    depth = orb.pipeline_depth()
    assert depth >= 2
    chunks = [(0, 2), (2, 4), (4, 5), (5, 7), (7, 8)]
    pinned = torch.from_numpy(seq).pin_memory().numpy()
    outs = []

    def collect(lo, hi):
        kps, desc, counts, good, ngood = orb.wait_batch()
        for i in range(hi - lo):
            f = lo + i
            assert_keypoints_equal(kps[i, :counts[i]], ext[f][0], "frame %d" % f)
            assert_descriptors_equal(desc[i, :counts[i]], ext[f][1], "frame %d" % f)
            if f == 0:
                assert ngood[i] == 0
            else:
                _check_matches(good[i], int(ngood[i]), matches[f])

    pending = []
    for lo, hi in chunks:
        if orb.batches_in_flight() == depth:
            collect(*pending.pop(0))
        orb.submit_batch(list(pinned[lo:hi]), m, 0.8, buffers(hi - lo))
        pending.append((lo, hi))
    assert orb.batches_in_flight() == depth
    with pytest.raises(Exception):              # the pipeline is full
        orb.submit_batch(list(pinned[0:1]), m, 0.8, buffers(1))
    # the blocking entry points refuse to run while batches are in flight
    with pytest.raises(Exception):
        orb.extract_batch(list(seq[:1]), cap=cap)
    while pending:
        collect(*pending.pop(0))
    assert orb.batches_in_flight() == 0
    with pytest.raises(Exception):
        orb.wait_batch()
    # the handle is usable through the blocking path again
    kps, desc, counts = orb.extract_batch(list(seq[:1]), cap=cap)
    assert_keypoints_equal(kps[0, :counts[0]], ext[0][0], "blocking after pipelined")
    m.close()
    orb.close()


def test_match_back_traverse():
    """orbx_match_back: every frame against its `back` predecessors (the reference's numBackTraverse loop,
    src/CameraPoseEstimator.cpp:405-409), across batch borders (history kept by the handle), back larger than a batch."""
    seq = syn.sequence(9, 640, 480, seed=24)
    nf, ratio, back = 600, 0.8, 3
    P = oracle.Params(nfeatures=nf)
    ext = [oracle.detect_and_compute(f, P) for f in seq]
    orb = ORB(nfeatures=nf, max_size=(640, 480), max_batch=4)
    m = BFMatcher()
    cap = orb.default_cap
    done = 0
    for chunk in (list(seq[:4]), list(seq[4:6]), list(seq[6:])):
        n = len(chunk)
        kps, desc, counts = orb.extract_batch(chunk, cap=cap)
        good, ngood = orb.match_back(m, back, ratio, cap, n)
        for i in range(n):
            f = done + i
            assert_descriptors_equal(desc[i, :counts[i]], ext[f][1], "frame %d" % f)
            for j in range(1, back + 1):
                if f - j < 0:
                    assert ngood[i, j - 1] == 0, "frame %d has no predecessor %d" % (f, j)
                else:
                    _check_matches(good[i, j - 1], int(ngood[i, j - 1]), oracle.match_features(ext[f][1], ext[f - j][1], ratio))
        done += n
    # a new sequence forgets the history
    orb.reset_sequence()
    kps, desc, counts = orb.extract_batch(list(seq[:2]), cap=cap)
    good, ngood = orb.match_back(m, 5, ratio, cap, 2)
    assert ngood[0].sum() == 0 and ngood[1, 1:].sum() == 0
    _check_matches(good[1, 0], int(ngood[1, 0]), oracle.match_features(ext[1][1], ext[0][1], ratio))
    m.close()
    orb.close()


def test_full_size_batch_consistency():
    """BASELINE configs[1] at full frame size: a 20-frame 1920x1080 batch (cut in two halves on two streams, uploaded in
    chunks) must equal the single-frame calls (graph replay) frame by frame, the pipelined path must equal the blocking
    one, and sampled frames must equal the oracle."""
    import torch
    seq = syn.sequence(20, 1920, 1080, seed=31)
    nf = 2000
    orb = ORB(nfeatures=nf, max_size=(1920, 1080), max_batch=20)
    m = BFMatcher()
    cap = orb.default_cap
    kps, desc, counts = orb.extract_batch(list(seq), cap=cap)
    good, ngood = orb.match_consecutive(m, 0.75, cap, 20)
    kps, desc, counts, good, ngood = kps.copy(), desc.copy(), counts.copy(), good.copy(), ngood.copy()
    P = oracle.Params(nfeatures=nf)
    for f in (0, 9, 10, 19):          # both halves, both ends
        ok, od = oracle.detect_and_compute(seq[f], P)
        assert_keypoints_equal(kps[f, :counts[f]], ok, "frame %d vs oracle" % f)
        assert_descriptors_equal(desc[f, :counts[f]], od, "frame %d vs oracle" % f)
    for f in range(20):
        k1, d1 = orb.detectAndCompute(seq[f])
        assert_keypoints_equal(k1, kps[f, :counts[f]], "single frame %d" % f)
        assert_descriptors_equal(d1, desc[f, :counts[f]], "single frame %d" % f)
    q, t, d = oracle.match_features(desc[10, :counts[10]], desc[9, :counts[9]], 0.75)    # across the split
    _check_matches(good[10], int(ngood[10]), (q, t, d))
    # pipelined path, same frames in two submissions
    orb.reset_sequence()
    outs = [(np.zeros((10, cap), KEYPOINT_DTYPE), np.zeros((10, cap, 32), np.uint8), np.zeros(10, np.int32),
             np.zeros((10, cap), DMATCH_DTYPE), np.zeros(10, np.int64)) for _ in range(2)]
    pinned = torch.from_numpy(seq).pin_memory().numpy()
    orb.submit_batch(list(pinned[:10]), m, 0.75, outs[0])
    orb.submit_batch(list(pinned[10:]), m, 0.75, outs[1])
    for b in range(2):
        k2, d2, c2, g2, n2 = orb.wait_batch()
        for i in range(10):
            f = 10 * b + i
            assert c2[i] == counts[f]
            assert np.array_equal(k2[i, :c2[i]], kps[f, :counts[f]]) and np.array_equal(d2[i, :c2[i]], desc[f, :counts[f]])
            assert n2[i] == ngood[f] and np.array_equal(g2[i, :n2[i]], good[f, :ngood[f]])
    m.close()
    orb.close()


def test_sharded_sequence_equals_single_rank():
    """ShardedSequence on the GPU: the blocks of a 3-rank split (run one after the other on this GPU) reproduce the
    single-rank run frame by frame, including the matches across block borders."""
    from monocular_slam_b200.sharded import ShardedSequence
    seq = syn.sequence(11, 640, 480, seed=41)
    orb = ORB(nfeatures=500, max_size=(640, 480), max_batch=3)
    m = BFMatcher()
    ref = {f: (k, d, g) for f, k, d, g in ShardedSequence(orb, m, 0.8, 0, 1).run(seq)}
    assert sorted(ref) == list(range(11)) and len(ref[0][2]) == 0
    P = oracle.Params(nfeatures=500)
    e4, e5 = oracle.detect_and_compute(seq[4], P), oracle.detect_and_compute(seq[5], P)
    _check_matches(ref[5][2], len(ref[5][2]), oracle.match_features(e5[1], e4[1], 0.8))
    got = {}
    for r in range(3):
        for f, k, d, g in ShardedSequence(orb, m, 0.8, r, 3).run(seq):
            assert f not in got
            got[f] = (k, d, g)
    assert sorted(got) == list(range(11))
    for f in range(11):
        assert np.array_equal(got[f][0], ref[f][0]) and np.array_equal(got[f][1], ref[f][1]) and np.array_equal(got[f][2], ref[f][2]), f
    m.close()
    orb.close()


def test_sharded_sequence_with_outlier_filter():
    """The filter rides along in ShardedSequence: per-frame status / F of a 2-rank split equal the single-rank run (the pair across
    the block border is computed by the rank that owns the later frame, from its own lead-in extraction)."""
    from monocular_slam_b200 import FundamentalFilter
    from monocular_slam_b200.sharded import ShardedSequence
    seq = syn.sequence(9, 640, 480, seed=43)
    orb = ORB(nfeatures=600, max_size=(640, 480), max_batch=4)
    m = BFMatcher()
    fm = FundamentalFilter()
    ref = {it[0]: it[1:] for it in ShardedSequence(orb, m, 0.8, 0, 1, fundamental=fm).run(seq)}
    assert sorted(ref) == list(range(9)) and len(ref[0][2]) == 0 and not ref[0][4].any()
    for f in range(1, 9):
        k, d, g, status, F = ref[f]
        Fh, sh, nh = fm.compute_fundamental(k, ref[f - 1][0], g)
        assert np.array_equal(status, sh) and np.array_equal(F, Fh)
    got = {}
    for r in range(2):
        for it in ShardedSequence(orb, m, 0.8, r, 2, fundamental=fm).run(seq):
            got[it[0]] = it[1:]
    assert sorted(got) == list(range(9))
    for f in range(9):
        for a, b in zip(got[f], ref[f]):
            assert np.array_equal(a, b), f
    fm.close()
    m.close()
    orb.close()


# ---------------------------------------------------------------------------------------------- boundary regressions
def test_sequence_buffers_follow_the_capacity_retry():
    """extract_batch retries with the maximum capacity after ORBX_E_CAPACITY; the sequence calls size their buffers from
    the batch the handle holds, and a stale (cap, nframes) is refused by Python and by the C ABI instead of overflowing."""
    import ctypes as C
    from monocular_slam_b200 import _lib
    seq = syn.sequence(3, 640, 480, seed=5)
    ext, matches = _oracle_sequence(seq, 500, 0.8)
    orb = ORB(nfeatures=500, max_size=(640, 480), max_batch=3)
    m = BFMatcher()
    kps, desc, counts = orb.extract_batch(list(seq), cap=100)          # too small: retried at max_keypoints
    assert kps.shape[1] == orb.max_keypoints
    with pytest.raises(ValueError):
        orb.match_consecutive(m, 0.8, 100, 3)                           # the cap the caller asked for is stale
    with pytest.raises(ValueError):
        orb.match_consecutive(m, 0.8, None, 2)
    good = np.zeros((3, 100), DMATCH_DTYPE)
    ngood = np.zeros(3, np.int64)
    rc = _lib.lib().orbx_match_consecutive(orb._h, m._h, 0.8, 3, 100, good.ctypes.data, ngood.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == _lib.E_INVALID and not good.view(np.uint8).any()
    good, ngood = orb.match_consecutive(m, 0.8)
    assert good.shape == (3, orb.max_keypoints)
    for f in (1, 2):
        _check_matches(good[f], int(ngood[f]), matches[f])
    # rows wider than the device rows: [n][cap] host layout, only the device's columns are written
    wide = np.zeros((3, orb.max_keypoints + 7), DMATCH_DTYPE)
    orb.extract_batch(list(seq), cap=orb.max_keypoints)
    orb.reset_sequence()
    rc = _lib.lib().orbx_match_consecutive(orb._h, m._h, 0.8, 3, orb.max_keypoints + 7, wide.ctypes.data,
                                           ngood.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0
    for f in (1, 2):
        _check_matches(wide[f], int(ngood[f]), matches[f])
    m.close()
    orb.close()


def test_sequence_calls_keep_the_callers_stream():
    """orbx_match_consecutive / orbx_filter_consecutive / orbx_submit_batch run the matcher and the filter on the
    extractor's stream for one call and must put back the stream the caller installed on those handles."""
    import ctypes as C
    import torch
    from monocular_slam_b200 import FundamentalFilter, _lib
    seq = syn.sequence(2, 640, 480, seed=6)
    orb = ORB(nfeatures=500, max_size=(640, 480), max_batch=2)
    m, fm = BFMatcher(), FundamentalFilter()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    m.set_stream(s1.cuda_stream)
    fm.set_stream(s2.cuda_stream)

    def streams():
        a, b = C.c_void_p(), C.c_void_p()
        _lib.check(_lib.lib().hamx_get_stream(m._h, C.byref(a)))
        _lib.check(_lib.lib().fmx_get_stream(fm._h, C.byref(b)))
        return a.value, b.value
    want = (s1.cuda_stream, s2.cuda_stream)
    assert streams() == want
    orb.extract_batch(list(seq))
    orb.match_consecutive(m, 0.8)
    orb.filter_consecutive(fm)
    assert streams() == want
    orb.extract_batch(list(seq))
    orb.match_back(m, 2, 0.8)
    orb.filter_back(fm, 2)
    assert streams() == want
    cap = orb.default_cap
    out = (np.zeros((2, cap), KEYPOINT_DTYPE), np.zeros((2, cap, 32), np.uint8), np.zeros(2, np.int32), np.zeros((2, cap), DMATCH_DTYPE),
           np.zeros(2, np.int64), np.zeros((2, cap), np.uint8), np.zeros((2, 3, 3), np.float64), np.zeros(2, np.int32))
    orb.submit_batch(list(seq), m, 0.8, out, fundamental=fm)
    orb.wait_batch()
    assert streams() == want
    m.close(); fm.close(); orb.close()


def test_pipeline_lanes_with_different_capacities():
    """Batches in flight with different capacities live in disjoint lane regions: results equal the blocking path's."""
    seq = syn.sequence(6, 640, 480, seed=8)
    ext, matches = _oracle_sequence(seq, 500, 0.8)
    orb = ORB(nfeatures=500, max_size=(640, 480), max_batch=2)
    m = BFMatcher()
    caps = [orb.default_cap, orb.max_keypoints, orb.default_cap + 64]
    outs = [(np.zeros((2, c), KEYPOINT_DTYPE), np.zeros((2, c, 32), np.uint8), np.zeros(2, np.int32), np.zeros((2, c), DMATCH_DTYPE),
             np.zeros(2, np.int64)) for c in caps]
    for b in range(3):
        orb.submit_batch(list(seq[2 * b:2 * b + 2]), m, 0.8, outs[b])
    for b in range(3):
        kps, desc, counts, good, ngood = orb.wait_batch()
        for i in range(2):
            f = 2 * b + i
            assert_keypoints_equal(kps[i, :counts[i]], ext[f][0], "frame %d" % f)
            assert_descriptors_equal(desc[i, :counts[i]], ext[f][1], "frame %d" % f)
            if i == 1:            # the link to the previous batch is dropped when the capacity changes (include/orbx.h)
                _check_matches(good[i], int(ngood[i]), matches[f])
    m.close()
    orb.close()


def test_check_dev_ignores_stale_overflow_flags():
    """An overflow left in a high frame slot by an earlier, larger _dev batch must not fail later, smaller batches."""
    import torch
    from monocular_slam_b200 import OrbxError
    seq = syn.sequence(4, 640, 480, seed=9)
    orb = ORB(nfeatures=500, max_size=(640, 480), max_batch=4)
    d_frames = torch.from_numpy(seq).cuda()
    cap = orb.default_cap
    kps = torch.zeros((4, cap, 7), dtype=torch.float32, device="cuda")
    desc = torch.zeros((4, cap, 32), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
    orb.extract_batch_dev(d_frames.data_ptr(), 640 * 480, 4, 640, 480, 640, kps.data_ptr(), desc.data_ptr(), 16, cnt.data_ptr())
    with pytest.raises(OrbxError):
        orb.check_dev()                     # 16 rows cannot hold 500 keypoints
    orb.extract_batch_dev(d_frames.data_ptr(), 640 * 480, 1, 640, 480, 640, kps.data_ptr(), desc.data_ptr(), cap, cnt.data_ptr())
    orb.check_dev()                         # slots 1..3 still carry the old flag; only slot 0 is in use
    orb.close()
