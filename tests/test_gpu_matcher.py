"""GPU parity: hamx_* (K6) through the C ABI against the oracle and the cv2 golden vectors -- bit-exact indices and
distances, ties broken by lowest train index (reference src/CameraPoseEstimator.cpp:200-213)."""
import os

import numpy as np
import pytest

import oracle
from helpers import knn_arrays
from monocular_slam_b200 import BFMatcher, DMATCH_DTYPE, TOP2_DTYPE, _lib, match_features
from monocular_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

MATCH = ["planted", "dup_rows", "zeros", "nt1", "nt2", "low_entropy", "ragged_33x65"]


KERNELS = {"auto": _lib.KERNEL_AUTO, "integer": _lib.KERNEL_INTEGER, "tensor": _lib.KERNEL_TENSOR}


@pytest.fixture(scope="module", params=list(KERNELS))
def matcher(request):
    """Every test below runs against the integer-pipe kernel (XOR + POPC), the tensor-core kernel (tcgen05 int8 contraction)
    and the size-based choice between them: the results must be the same bits."""
    m = BFMatcher()
    m.set_kernel(KERNELS[request.param])
    yield m
    m.close()


def _check(matcher, q, t):
    m, c = matcher.knnMatch(q, t, 2)
    idx, dist = knn_arrays(m, c)
    oi, od = oracle.knn2(q, t)
    assert np.array_equal(idx, oi), "train indices differ"
    assert np.array_equal(dist, od), "distances differ"
    assert (m["query_idx"] == np.arange(len(q))[:, None]).all() and (m["img_idx"] == 0).all()
    assert (c == min(len(t), 2)).all()
    return oi, od


@pytest.mark.parametrize("name", MATCH)
def test_golden_cases(matcher, golden_dir, name):
    g = np.load(os.path.join(golden_dir, "matcher_cases.npz"))
    q, t = g[f"{name}_q"], g[f"{name}_t"]
    m, c = matcher.knnMatch(q, t, 2)
    idx, dist = knn_arrays(m, c)
    assert np.array_equal(idx, g[f"{name}_idx"]) and np.array_equal(dist, g[f"{name}_dist"])
    if len(t) >= 2:
        for r in (0.75, 0.8, 0.85):
            good = matcher.match_ratio(q, t, r)
            got = np.stack([good["query_idx"], good["train_idx"], good["distance"].astype(np.int32)], 1).reshape(-1, 3)
            assert np.array_equal(got, g[f"{name}_good_{int(r * 100)}"]), "ratio %.2f" % r


@pytest.mark.parametrize("score", ["harris", "fast"])
def test_golden_frame_pair(matcher, golden_dir, score):
    g = np.load(os.path.join(golden_dir, "kitti_pair.npz"))
    m, c = matcher.knnMatch(g[f"{score}_desc0"], g[f"{score}_desc1"], 2)
    idx, dist = knn_arrays(m, c)
    assert np.array_equal(idx, g[f"{score}_knn_idx"]) and np.array_equal(dist, g[f"{score}_knn_dist"])
    good = match_features(g[f"{score}_desc0"], g[f"{score}_desc1"], 0.75)
    got = np.stack([good["query_idx"], good["train_idx"], good["distance"].astype(np.int32)], 1)
    assert np.array_equal(got, g[f"{score}_good_75"])


@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 2), (5, 3), (127, 128), (128, 129), (257, 1025), (2000, 2000), (2115, 2064),
                                   (300, 70000), (40000, 513), (8000, 8000)])
def test_random_sizes(matcher, nq, nt):
    _check(matcher, syn.descriptors(nq * 7 + 1, nq), syn.descriptors(nt * 3 + 2, nt))


def test_heavy_ties(matcher):
    rng = np.random.default_rng(3)
    base = (rng.integers(0, 2, (8, 32)) * 255).astype(np.uint8)
    t = base[rng.integers(0, 8, 5000)]
    q = base[rng.integers(0, 8, 700)]
    _check(matcher, q, t)


def test_empty(matcher):
    m, c = matcher.knnMatch(np.zeros((0, 32), np.uint8), syn.descriptors(1, 10), 2)
    assert m.shape == (0, 2)
    m, c = matcher.knnMatch(syn.descriptors(1, 10), np.zeros((0, 32), np.uint8), 2)
    assert (c == 0).all() and (m["train_idx"] == -1).all()
    assert len(matcher.match_ratio(syn.descriptors(1, 10), np.zeros((0, 32), np.uint8), 0.8)) == 0
    assert len(matcher.match_ratio(syn.descriptors(1, 10), syn.descriptors(2, 1), 0.8)) == 0


def test_rejects_bad_shapes(matcher):
    with pytest.raises(ValueError):
        matcher.knnMatch(np.zeros((4, 16), np.uint8), np.zeros((4, 32), np.uint8), 2)
    with pytest.raises(ValueError):
        BFMatcher(normType=4)


def test_sharded_merge_equals_single(matcher):
    """Train set cut into shards, per-shard top-2 with global offsets, merged: identical to one pass."""
    import torch
    t = syn.descriptors(11, 20000)
    t[5000:5010] = t[100:110]        # duplicates across shard boundaries exercise the tie rule in the merge
    q = syn.planted_queries(12, t, 3000)
    oi, od = oracle.knn2(q, t)
    dq = torch.from_numpy(q).cuda()
    nparts = 8
    bounds = np.linspace(0, len(t), nparts + 1).astype(int)
    parts = torch.empty((nparts, len(q), 4), dtype=torch.int32, device="cuda")
    shards = [torch.from_numpy(t[bounds[i]:bounds[i + 1]]).cuda() for i in range(nparts)]
    matcher.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        for i in range(nparts):
            matcher.knn2_dev(dq.data_ptr(), len(q), shards[i].data_ptr(), len(shards[i]), int(bounds[i]), parts[i].data_ptr())
        out = torch.empty((len(q), 4), dtype=torch.int32, device="cuda")
        matcher.merge_top2_dev(parts.data_ptr(), nparts, len(q), out.data_ptr())
        good = torch.empty((len(q), 4), dtype=torch.int32, device="cuda")
        ngood = torch.zeros(1, dtype=torch.int64, device="cuda")
        matcher.ratio_dev(out.data_ptr(), len(q), 0.75, good.data_ptr(), ngood.data_ptr())
        torch.cuda.synchronize()
    finally:
        matcher.set_stream(None)
    r = out.cpu().numpy()
    assert np.array_equal(r[:, [1, 3]], oi) and np.array_equal(r[:, [0, 2]], od)
    gq, gt, gd = oracle.ratio_test(oi, od, 0.75)
    n = int(ngood.item())
    gm = good.cpu().numpy()[:n].copy().view(DMATCH_DTYPE).reshape(-1)
    assert n == len(gq) and np.array_equal(gm["query_idx"], gq) and np.array_equal(gm["train_idx"], gt)
    assert np.array_equal(gm["distance"].astype(np.int32), gd)


def test_large_sampled(matcher):
    """200k x 2000 (BASELINE config 4 shape): all queries checked against the oracle on the full train set."""
    t = syn.descriptors(4, 200000)
    q = syn.planted_queries(5, t, 2000)
    _check(matcher, q, t)


def test_large_properties(matcher):
    """A size the oracle cannot sweep in seconds: 60k x 300k.  Checked through properties: sampled rows against the
    oracle, d0 <= d1, indices in range, and invariance of the result to reversing the query order."""
    t = syn.descriptors(6, 300000)
    q = syn.descriptors(7, 60000)
    m, c = matcher.knnMatch(q, t, 2)
    idx, dist = knn_arrays(m, c)
    assert (dist[:, 0] <= dist[:, 1]).all() and (idx >= 0).all() and (idx < len(t)).all()
    assert ((dist[:, 0] < dist[:, 1]) | (idx[:, 0] < idx[:, 1])).all()
    sample = np.random.default_rng(0).choice(len(q), 200, replace=False)
    oi, od = oracle.knn2(q[sample], t)
    assert np.array_equal(idx[sample], oi) and np.array_equal(dist[sample], od)
    m2, c2 = matcher.knnMatch(q[::-1], t, 2)
    i2, d2 = knn_arrays(m2, c2)
    assert np.array_equal(i2[::-1], idx) and np.array_equal(d2[::-1], dist)


@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("world,nq,nt", [(1, 300, 1000), (2, 2000, 5001), (3, 777, 130), (4, 257, 4096), (8, 2000, 200000)])
def test_p2p_fused_scatter_merge_single_process(world, nq, nt, kernel):
    """hamx_knn2_p2p_*: `world` ranks emulated by `world` handles of one process on one GPU (same-process pointer import).
    All scatters are queued before any merge, so no kernel ever waits on a later launch.  Twice, to cover both buffer
    parities and the reuse of the completion counter."""
    import torch
    from monocular_slam_b200.sharded import shard_bounds
    q = syn.descriptors(31, nq)
    t = syn.descriptors(32, nt)
    t[nt // 2] = q[5]                    # an exact hit and duplicated rows across shard borders: ties by lowest index
    t[::97] = t[0]
    want_idx, want_dist = oracle.knn2(q, t)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ms = [BFMatcher() for _ in range(world)]
        for m in ms:
            m.set_stream(stream.cuda_stream)
            m.set_kernel(KERNELS[kernel])
        bases = [m.p2p_export(nq + 13, world, r)[1] for r, m in enumerate(ms)]
        for m in ms:
            m.p2p_import_ptrs(bases)
        dq = torch.from_numpy(q).cuda()
        dt = torch.from_numpy(t).cuda()
        b = shard_bounds(nt, world)
        for rep in range(3):
            outs = [torch.full((nq, 4), -7, dtype=torch.int32, device="cuda") for _ in range(world)]
            for r, m in enumerate(ms):
                lo, hi = int(b[r]), int(b[r + 1])
                shard = dt[lo:hi].contiguous() if hi > lo else torch.zeros((1, 32), dtype=torch.uint8, device="cuda")
                m.knn2_p2p_scatter_dev(dq.data_ptr(), nq, shard.data_ptr(), hi - lo, lo)
            for r, m in enumerate(ms):
                m.p2p_merge_dev(nq, outs[r].data_ptr())
            stream.synchronize()
            for r in range(world):
                got = outs[r].cpu().numpy()
                assert np.array_equal(got[:, 1], want_idx[:, 0]) and np.array_equal(got[:, 3], want_idx[:, 1]), "rank %d rep %d" % (r, rep)
                assert np.array_equal(got[:, 0], want_dist[:, 0]) and np.array_equal(got[:, 2], want_dist[:, 1])
        for m in ms:
            m.close()


@pytest.mark.parametrize("kernel", ["integer", "tensor"])
def test_train_set_larger_than_index_field(kernel):
    """More than 2^23 train rows: the packed (distance << 23 | index) key covers one chunk, the host splits the train set and
    merges the chunks with the 64-bit rule.  Planted exact duplicates on both sides of the chunk border pin the tie rule
    (lowest global index first) and the global offsets; a sample of queries is checked against the oracle on a window."""
    import torch
    nt, nq = (1 << 23) + 70001, 512
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device="cuda", generator=g)
        q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda", generator=g)
        border = 1 << 23
        # query i (< 64) has two exact copies: one just before the border, one after it
        for i in range(64):
            t[border - 100 + i] = q[i]
            t[border + 5000 + i] = q[i]
        # query 64 + i (< 64): copies only in the second chunk, in descending position
        for i in range(64):
            t[border + 60000 - i] = q[64 + i]
        m = BFMatcher()
        m.set_stream(stream.cuda_stream)
        m.set_kernel(KERNELS[kernel])
        out = torch.empty((nq, 4), dtype=torch.int32, device="cuda")
        m.knn2_dev(q.data_ptr(), nq, t.data_ptr(), nt, 0, out.data_ptr())
        stream.synchronize()
    r = out.cpu().numpy()
    for i in range(64):
        assert tuple(r[i]) == (0, border - 100 + i, 0, border + 5000 + i), (i, r[i])
        assert r[64 + i][0] == 0 and r[64 + i][1] == border + 60000 - i
    # the rest: compare with the oracle on the rows around the reported neighbours plus a random window (distances must be
    # true distances and no closer row may exist in the window)
    th = t.cpu().numpy()
    qh = q.cpu().numpy()
    win = np.concatenate([np.arange(0, 200000), np.arange(border - 100000, border + 70001)])
    oi, od = oracle.knn2(qh[128:192], th[win])
    for j in range(64):
        i = 128 + j
        d0 = int(np.unpackbits(qh[i] ^ th[r[i][1]]).sum())
        d1 = int(np.unpackbits(qh[i] ^ th[r[i][3]]).sum())
        assert (d0, d1) == (r[i][0], r[i][2]) and d0 <= d1
        assert d0 <= od[j, 0] and d1 <= od[j, 1]          # nothing in the window beats the global answer
    m.close()


def test_kernels_agree_on_the_benchmark_shape():
    """1 M x 125 k per GPU (BASELINE config 5): far beyond the oracle; the two kernels, which share no arithmetic (XOR + POPC
    vs a +-1 int8 contraction), must produce identical top-2 for every query, and a sample must equal the oracle."""
    import torch
    nq, nt = 1 << 18, 125000
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        g = torch.Generator(device="cuda"); g.manual_seed(9)
        q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda", generator=g)
        t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device="cuda", generator=g)
        t[::1013] = t[7]                                  # equal rows: ties by lowest index, in both kernels
        outs = []
        for mode in ("integer", "tensor"):
            m = BFMatcher()
            m.set_stream(stream.cuda_stream)
            m.set_kernel(KERNELS[mode])
            o = torch.empty((nq, 4), dtype=torch.int32, device="cuda")
            m.knn2_dev(q.data_ptr(), nq, t.data_ptr(), nt, 3, o.data_ptr())
            stream.synchronize()
            outs.append(o)
            m.close()
    assert torch.equal(outs[0], outs[1])
    sample = np.random.default_rng(1).choice(nq, 64, replace=False)
    oi, od = oracle.knn2(q.cpu().numpy()[sample], t.cpu().numpy())
    r = outs[1].cpu().numpy()[sample]
    assert np.array_equal(r[:, [1, 3]], oi + 3) and np.array_equal(r[:, [0, 2]], od)
