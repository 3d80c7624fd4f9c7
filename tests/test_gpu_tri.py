"""GPU parity: triangulation and map-point association (csrc/triangulate.cu, include/orbx.h "trx_*") against the oracle
(oracle/tri_oracle.c) and the committed cv2 goldens (tests/golden/tri_cases.npz, made by make_golden_tri.py).

Tolerances: X within 1e-9 relative of cv2 / the oracle (double-precision SVD of a 4x4 system by two implementations);
front-of-camera flags, their counts, the winning hypothesis, association lists and accept masks identical."""
import os

import numpy as np
import pytest

import oracle
from monocular_slam_b200 import CAMERAS_DTYPE, DMATCH_DTYPE, KEYPOINT_DTYPE, Triangulator
from test_oracle_tri import CASES, G, random_lists

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rel_err(X, ref):
    n = np.linalg.norm(ref, axis=1)
    return float((np.linalg.norm(X - ref, axis=1) / np.where(n > 0, n, 1)).max()) if len(ref) else 0.0


@pytest.fixture(scope="module")
def tri():
    t = Triangulator()
    yield t
    t.close()


@pytest.mark.parametrize("name", CASES)
def test_triangulate_matches_cv2_golden(tri, name):
    g = {k.split("/")[1]: G[k] for k in G.files if k.startswith(name + "/")}
    X, front, count = tri.triangulate(g["p1"], g["p2"], g["Rt1"], g["Rt2"], g["K1"], g["K2"])
    assert np.array_equal(front, g["front"]) and count == int(g["front"].sum())
    assert rel_err(X, g["X"]) <= TOL and rel_err(X, g["Xtp"]) <= TOL
    res, cnt = tri.TriangulateMultiplePointsFromTwoView(g["p1"], g["p2"], g["Rt1"], g["Rt2"], g["K1"], g["K2"], countFront=True)
    assert cnt == count and np.array_equal(res, X)
    assert tri.TriangulateMultiplePointsFromTwoView(g["p1"], g["p2"], g["Rt1"], g["Rt2"], g["K1"], g["K2"])[1] == 0
    if "hyp_Rts" in g:
        best, counts, Xb = tri.triangulate_hypotheses(g["p1"], g["p2"], g["Rt1"], g["hyp_Rts"], g["K1"], g["K2"])
        assert best == int(g["hyp_best"]) and np.array_equal(counts, g["hyp_counts"])
        ob, oc, oX = oracle.triangulate_best(g["p1"], g["p2"], g["Rt1"], g["hyp_Rts"], g["K1"], g["K2"])
        assert best == ob and rel_err(Xb, oX) <= TOL


def test_triangulate_empty_and_errors(tri):
    from monocular_slam_b200 import OrbxError
    I = np.c_[np.eye(3), np.zeros(3)]
    X, front, n = tri.triangulate(np.zeros((0, 2)), np.zeros((0, 2)), I, I, np.eye(3), np.eye(3))
    assert X.shape == (0, 3) and n == 0
    with pytest.raises(ValueError):
        tri.triangulate(np.zeros((3, 2)), np.zeros((3, 2)), np.eye(3), I, np.eye(3), np.eye(3))
    best, counts, _ = tri.triangulate_hypotheses(np.zeros((0, 2)), np.zeros((0, 2)), I, np.stack([I, I]), np.eye(3), np.eye(3))
    assert best == 0 and not counts.any()
    with pytest.raises(OrbxError):
        tri.triangulate_batch(np.zeros((1, 4, 2), np.float32), np.zeros((1, 4, 2), np.float32), [4], np.zeros((1, 65), CAMERAS_DTYPE))


def _cams(Rt1, Rt2, K1, K2):
    c = np.zeros((), CAMERAS_DTYPE)
    c["Rt1"], c["Rt2"], c["K1"], c["K2"] = Rt1, Rt2, K1, K2
    return c


def test_triangulate_batch_with_select_and_hypotheses(tri):
    """Ragged problems, a select mask, several hypotheses per problem -- each (problem, hypothesis) against the oracle."""
    names = [n for n in CASES if "hyp_Rts" in [k.split("/")[1] for k in G.files if k.startswith(n + "/")]]
    cap = max(len(G[n + "/p1"]) for n in names) + 5
    nprob = len(names)
    p1 = np.zeros((nprob, cap, 2), np.float32)
    p2 = np.zeros((nprob, cap, 2), np.float32)
    counts = np.zeros(nprob, np.int32)
    cams = np.zeros((nprob, 4), CAMERAS_DTYPE)
    r = np.random.default_rng(3)
    select = (r.random((nprob, cap)) < 0.7).astype(np.uint8)
    for i, n in enumerate(names):
        k = len(G[n + "/p1"])
        counts[i] = k
        p1[i, :k], p2[i, :k] = G[n + "/p1"], G[n + "/p2"]          # float32-representable by construction
        for h in range(4):
            cams[i, h] = _cams(G[n + "/Rt1"], G[n + "/hyp_Rts"][h], G[n + "/K1"], G[n + "/K2"])
    X, front, nfront, best = tri.triangulate_batch(p1, p2, counts, cams, select=select, want_best=True)
    for i, n in enumerate(names):
        k = int(counts[i])
        sel = select[i, :k] > 0
        for h in range(4):
            oX, ofr, _ = oracle.triangulate(G[n + "/p1"], G[n + "/p2"], G[n + "/Rt1"], G[n + "/hyp_Rts"][h], G[n + "/K1"], G[n + "/K2"])
            assert np.array_equal(front[i, h, :k][sel], ofr[sel]) and not front[i, h, :k][~sel].any() and not front[i, h, k:].any()
            assert rel_err(X[i, h, :k][sel], oX[sel]) <= TOL and not X[i, h, :k][~sel].any() and not X[i, h, k:].any()
            assert nfront[i, h] == int(ofr[sel].sum())
        assert best[i] == int(np.argmax(nfront[i]))
    X1, front1, nfront1 = tri.triangulate_batch(p1, p2, counts, cams[:, 0])
    assert X1.shape == (nprob, 1, cap, 3) and nfront1[0, 0] == int(oracle.triangulate(
        G[names[0] + "/p1"], G[names[0] + "/p2"], G[names[0] + "/Rt1"], G[names[0] + "/hyp_Rts"][0], G[names[0] + "/K1"], G[names[0] + "/K2"])[2])


@pytest.mark.parametrize("seed,nprob,back,cap,ncur,npre,frac,with_status", [
    (1, 3, 5, 64, 50, 60, 0.5, False), (2, 2, 5, 2500, 2300, 2400, 0.3, True), (3, 4, 1, 40, 40, 10, 0.9, False),
    (4, 1, 3, 700, 20, 700, 1.0, True), (5, 2, 5, 32, 30, 30, 0.0, False), (6, 3, 8, 600, 500, 40, 0.6, True),
    (7, 1, 5, 2500, 2500, 2500, 0.05, False)])
def test_association_and_new_point_selection(tri, seed, nprob, back, cap, ncur, npre, frac, with_status):
    """trx_associate_dev / trx_select_new_dev == the reference's two loops (oracle.associate / select_new), list for list."""
    import torch
    good = np.zeros((nprob, back, cap), DMATCH_DTYPE)
    ngood = np.zeros((nprob, back), np.int64)
    premap = np.zeros((nprob, back, cap), np.int32)
    status = np.ones((nprob, back, cap), np.uint8)
    r = np.random.default_rng(seed)
    for p in range(nprob):
        m, n, pm = random_lists(100 * seed + p, back, cap, ncur, npre, frac)
        good[p], ngood[p], premap[p] = m.view(DMATCH_DTYPE), n, pm
    if with_status:
        status = (r.random(status.shape) < 0.8).astype(np.uint8)
    dev = "cuda"
    d_good = torch.from_numpy(good.view(np.int32).reshape(nprob, back, cap, 4)).to(dev)
    d_ngood = torch.from_numpy(ngood).to(dev)
    d_status = torch.from_numpy(status).to(dev)
    d_premap = torch.from_numpy(premap).to(dev)
    d_ncur = torch.full((nprob,), ncur, dtype=torch.int32, device=dev)
    d_cur = torch.zeros((nprob, cap), dtype=torch.int32, device=dev)
    d_aq = torch.zeros((nprob, cap), dtype=torch.int32, device=dev)
    d_amp = torch.zeros((nprob, cap), dtype=torch.int32, device=dev)
    d_na = torch.zeros(nprob, dtype=torch.int32, device=dev)
    st_ptr = d_status.data_ptr() if with_status else 0
    tri.associate_dev(d_good.data_ptr(), d_ngood.data_ptr(), st_ptr, d_premap.data_ptr(), d_ncur.data_ptr(), nprob, back, cap,
                      d_cur.data_ptr(), d_aq.data_ptr(), d_amp.data_ptr(), d_na.data_ptr())
    d_next = torch.arange(nprob, dtype=torch.int32, device=dev) * 100000 + 7
    d_acc = torch.zeros((nprob, back, cap), dtype=torch.uint8, device=dev)
    d_nnew = torch.zeros(nprob, dtype=torch.int32, device=dev)
    d_premap2 = d_premap.clone()
    d_cur2 = d_cur.clone()
    tri.select_new_dev(d_good.data_ptr(), d_ngood.data_ptr(), st_ptr, d_premap2.data_ptr(), d_cur2.data_ptr(), d_ncur.data_ptr(),
                       d_next.data_ptr(), nprob, back, cap, d_acc.data_ptr(), d_nnew.data_ptr())
    tri.synchronize()
    cur, aq, amp, na = d_cur.cpu().numpy(), d_aq.cpu().numpy(), d_amp.cpu().numpy(), d_na.cpu().numpy()
    acc, nnew, premap2, cur2 = d_acc.cpu().numpy(), d_nnew.cpu().numpy(), d_premap2.cpu().numpy(), d_cur2.cpu().numpy()
    total_new = 0
    for p in range(nprob):
        # the reference compacts every list to its inliers before the loops (:417-421)
        lists = np.zeros((back, cap), DMATCH_DTYPE)
        cnt = np.zeros(back, np.int32)
        pos = []
        for l in range(back):
            keep = np.nonzero(status[p, l, :ngood[p, l]])[0]
            cnt[l] = len(keep)
            lists[l, :len(keep)] = good[p, l, keep]
            pos.append(keep)
        wc, wq, wm = oracle.associate(lists, cnt, premap[p], ncur)
        assert na[p] == len(wq) and np.array_equal(aq[p, :na[p]], wq) and np.array_equal(amp[p, :na[p]], wm)
        assert np.array_equal(cur[p, :ncur], wc) and (cur[p, ncur:] == -1).all()
        wa, wp, wcm, wk = oracle.select_new(lists, cnt, premap[p], np.r_[wc, np.full(cap - ncur, -1, np.int32)], next_id=100000 * p + 7)
        assert nnew[p] == wk
        for l in range(back):
            full = np.zeros(cap, np.uint8)
            full[pos[l]] = wa[l, :cnt[l]]
            assert np.array_equal(acc[p, l], full), "problem %d list %d" % (p, l)
        assert np.array_equal(premap2[p], wp) and np.array_equal(cur2[p], wcm)
        total_new += wk
    if frac < 1.0 and ncur > 10:
        assert total_new > 0


def test_device_chain_filter_to_triangulation(tri):
    """hamx_match_back_dev -> fmx_filter_back_dev -> trx_select_new_dev -> trx_triangulate_back_dev with nothing but device
    pointers in between; every stage after the filter is checked against the oracle fed with the downloaded lists."""
    import torch
    from monocular_slam_b200 import ORB, BFMatcher, FundamentalFilter
    from monocular_slam_b200 import synthetic as syn
    dev = "cuda"
    frames = syn.layered_views(4, 640, 480, 3)
    n, back, W, H = 3, 2, 640, 480
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        orb = ORB(nfeatures=600, max_size=(W, H), max_batch=n)
        bf, fm = BFMatcher(), FundamentalFilter()
        for o in (orb, bf, fm, tri):
            o.set_stream(stream.cuda_stream)
        cap = orb.default_cap
        d_fr = torch.from_numpy(frames).to(dev)
        d_kps = torch.zeros((n, cap, 7), dtype=torch.float32, device=dev)
        d_desc = torch.zeros((n, cap, 32), dtype=torch.uint8, device=dev)
        d_cnt = torch.zeros(n, dtype=torch.int32, device=dev)
        orb.extract_batch_dev(d_fr.data_ptr(), W * H, n, W, H, W, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_cnt.data_ptr())
        np_ = n * back
        d_good = torch.zeros((np_, cap, 4), dtype=torch.int32, device=dev)
        d_ngood = torch.zeros(np_, dtype=torch.int64, device=dev)
        from monocular_slam_b200 import _lib
        _lib.check(_lib.lib().hamx_match_back_dev(bf._h, d_desc.data_ptr(), d_cnt.data_ptr(), n, cap, back, None, None, 0, 0.8,
                                                  d_good.data_ptr(), d_ngood.data_ptr()))
        d_status = torch.zeros((np_, cap), dtype=torch.uint8, device=dev)
        d_F = torch.zeros((np_, 9), dtype=torch.float64, device=dev)
        d_info = torch.zeros((np_, 4), dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().fmx_filter_back_dev(fm._h, d_kps.data_ptr(), n, cap, back, None, 0, d_good.data_ptr(), d_ngood.data_ptr(), 3.0, 0.85,
                                                  d_status.data_ptr(), d_F.data_ptr(), d_info.data_ptr()))
        # every frame is its own association problem; nothing has a map point yet, so every inlier is a candidate
        d_premap = torch.full((n, back, cap), -1, dtype=torch.int32, device=dev)
        d_cur = torch.full((n, cap), -1, dtype=torch.int32, device=dev)
        d_acc = torch.zeros((n, back, cap), dtype=torch.uint8, device=dev)
        d_nnew = torch.zeros(n, dtype=torch.int32, device=dev)
        tri.select_new_dev(d_good.data_ptr(), d_ngood.data_ptr(), d_status.data_ptr(), d_premap.data_ptr(), d_cur.data_ptr(), d_cnt.data_ptr(), 0,
                           n, back, cap, d_acc.data_ptr(), d_nnew.data_ptr())
        K = np.array([[520.0, 0, 320], [0, 520.0, 240], [0, 0, 1]])
        poses = [np.c_[np.eye(3), np.zeros(3)], np.c_[np.eye(3), np.array([-0.1, -0.04, 0.0])], np.c_[np.eye(3), np.array([-0.2, -0.08, 0.0])]]
        cams = np.zeros((n, back), CAMERAS_DTYPE)
        for f in range(n):
            for j in range(1, back + 1):
                if f - j >= 0:
                    cams[f, j - 1]["Rt1"], cams[f, j - 1]["Rt2"] = poses[f - j], poses[f]      # view 1 = the predecessor (:504-506)
                cams[f, j - 1]["K1"] = cams[f, j - 1]["K2"] = K
        d_cams = torch.from_numpy(cams.view(np.float64).reshape(np_, 42)).to(dev)
        d_X = torch.zeros((np_, cap, 3), dtype=torch.float64, device=dev)
        d_front = torch.zeros((np_, cap), dtype=torch.uint8, device=dev)
        d_nfront = torch.zeros(np_, dtype=torch.int32, device=dev)
        tri.triangulate_back_dev(d_kps.data_ptr(), n, cap, back, 0, 0, d_good.data_ptr(), d_ngood.data_ptr(), d_acc.data_ptr(), d_cams.data_ptr(),
                                 d_X.data_ptr(), d_front.data_ptr(), d_nfront.data_ptr())
    stream.synchronize()
    kps = d_kps.cpu().numpy().view(KEYPOINT_DTYPE).reshape(n, cap)
    good = d_good.cpu().numpy().view(DMATCH_DTYPE).reshape(n, back, cap)
    ngood, status = d_ngood.cpu().numpy().reshape(n, back), d_status.cpu().numpy().reshape(n, back, cap)
    acc, X, front, nfront = d_acc.cpu().numpy(), d_X.cpu().numpy().reshape(n, back, cap, 3), d_front.cpu().numpy().reshape(n, back, cap), d_nfront.cpu().numpy().reshape(n, back)
    assert ngood[1, 0] > 100 and ngood[2, 0] > 100 and ngood[2, 1] > 50 and ngood[0].sum() == 0
    checked = 0
    for f in range(n):
        lists = np.zeros((back, cap), DMATCH_DTYPE)
        cnt = np.zeros(back, np.int32)
        pos = []
        for l in range(back):
            keep = np.nonzero(status[f, l, :ngood[f, l]])[0]
            cnt[l] = len(keep)
            lists[l, :len(keep)] = good[f, l, keep]
            pos.append(keep)
        wa, _, _, wk = oracle.select_new(lists, cnt, np.full((back, cap), -1, np.int32), np.full(cap, -1, np.int32))
        for l in range(back):
            full = np.zeros(cap, np.uint8)
            full[pos[l]] = wa[l, :cnt[l]]
            assert np.array_equal(acc[f, l], full)
            sel = np.nonzero(full)[0]
            if f - (l + 1) < 0:
                assert not len(sel)
                continue
            m = good[f, l, sel]
            pre, cur = kps[f - l - 1][m["train_idx"]], kps[f][m["query_idx"]]
            p1 = np.stack([pre["x"], pre["y"]], 1).astype(np.float64)
            p2 = np.stack([cur["x"], cur["y"]], 1).astype(np.float64)
            oX, ofr, ocnt = oracle.triangulate(p1, p2, poses[f - l - 1], poses[f], K, K)
            assert np.array_equal(front[f, l, sel], ofr) and nfront[f, l] == ocnt
            assert rel_err(X[f, l, sel], oX) <= 1e-7       # near-degenerate rays appear among real matches: looser than the goldens
            unsel = np.setdiff1d(np.arange(cap), sel)
            assert not X[f, l, unsel].any() and not front[f, l, unsel].any()
            checked += len(sel)
    assert checked > 200
    bf.close(); fm.close(); orb.close()
