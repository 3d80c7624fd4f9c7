"""GPU parity: loop-closure candidate scoring (hamx_nbest*, hamx_loop_score*) against the oracle's literal restatement of
LoopCloser::NBestMatches / DetectLoop (oracle/loop_oracle.c; reference src/LoopCloser.cpp:19-105).  Integer work: everything
must be identical, including the reference's order among equal distances."""
import numpy as np
import pytest

import oracle
from monocular_slam_b200 import BFMatcher, _lib
from monocular_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["auto", "integer", "tensor"])
def bf(request):
    """The scoring runs on the integer pipes (XOR + POPC) or on the tensor cores (the matcher's int8 contraction with a
    counting epilogue); every test must hold for both and for the size-based choice."""
    m = BFMatcher()
    m.set_kernel({"auto": _lib.KERNEL_AUTO, "integer": _lib.KERNEL_INTEGER, "tensor": _lib.KERNEL_TENSOR}[request.param])
    yield m
    m.close()


@pytest.mark.parametrize("nq,nt,n", [(300, 700, 10), (1, 1, 10), (17, 5, 10), (0, 50, 10), (40, 0, 3), (257, 129, 1), (500, 2000, 16),
                                      (129, 1000, 4), (2000, 2000, 10)])
def test_nbest_matches_oracle(bf, nq, nt, n):
    t = syn.descriptors(100 + nt, nt)
    q = syn.planted_queries(200 + nq, t, nq) if nt and nq else syn.descriptors(7, nq)
    if nt > 40:
        t[20:30] = t[5]                  # ten identical rows: the tie order of the reference's insertion is exercised
        t[35] = t[5]
        if nq > 3:
            q[0] = t[5]
            q[1] = t[5]
            q[1, 3] ^= 0x10
    d, i = bf.NBestMatches(q, t, n)
    wd, wi = oracle.nbest(q, t, n)
    assert np.array_equal(d, wd) and np.array_equal(i, wi)


def test_nbest_low_entropy_ties(bf):
    """Descriptors drawn from 4 byte values only: most distances collide, so the result hinges on the tie rule."""
    r = np.random.default_rng(3)
    t = r.choice(np.array([0, 255, 15, 240], np.uint8), (600, 32))
    q = r.choice(np.array([0, 255, 15, 240], np.uint8), (150, 32))
    d, i = bf.NBestMatches(q, t, 10)
    wd, wi = oracle.nbest(q, t, 10)
    assert np.array_equal(d, wd) and np.array_equal(i, wi)


@pytest.mark.parametrize("nq,nf,cap,n,thr", [(200, 9, 300, 10, 60), (2000, 12, 2100, 10, 40), (1, 3, 10, 10, 256), (300, 4, 128, 3, 118),
                                               (257, 5, 256, 16, 100), (700, 333, 129, 10, 125), (513, 40, 1000, 2, 131)])
def test_loop_score_matches_oracle(bf, nq, nf, cap, n, thr):
    r = np.random.default_rng(nq + nf)
    q = r.integers(0, 256, (nq, 32), dtype=np.uint8)
    frames = r.integers(0, 256, (nf, cap, 32), dtype=np.uint8)
    counts = r.integers(0, cap + 1, nf).astype(np.int32)
    counts[0] = cap
    if nf > 2:
        counts[1] = 0
    # two frames that really share content with the query frame (noisy copies), the later one slightly better
    for f, flips in ((nf - 1, 12), (nf // 2, 10)):
        k = min(nq, counts[f], cap)
        if k:
            frames[f, :k] = q[:k]
            for row in range(k):
                bits = r.choice(256, flips, replace=False)
                for b in bits:
                    frames[f, row, b >> 3] ^= np.uint8(1 << (b & 7))
    scores, best = bf.loop_score(q, frames, counts, n, thr)
    want, wb = oracle.loop_score(q, frames, counts, n, thr)
    assert np.array_equal(scores, want) and best == wb


def test_loop_score_no_candidate_and_first_maximum(bf):
    r = np.random.default_rng(1)
    q = r.integers(0, 256, (64, 32), dtype=np.uint8)
    frames = r.integers(0, 256, (5, 64, 32), dtype=np.uint8)
    counts = np.full(5, 64, np.int32)
    s, b = bf.loop_score(q, frames, counts, 10, 20)          # random 256-bit strings are ~128 bits apart
    assert not s.any() and b == -1
    frames[1] = q
    frames[4] = q                                            # equal scores: the first frame wins (strict '>', :42)
    s, b = bf.loop_score(q, frames, counts, 10, 20)
    assert s[1] == s[4] == 64 and b == 1
    s, b = bf.loop_score(q, frames[:0], counts[:0], 10, 20)
    assert len(s) == 0 and b == -1


def test_loop_score_on_orb_descriptors(bf):
    """Real descriptors: a revisited view must outscore unrelated frames, and the scores equal the oracle's."""
    from monocular_slam_b200 import ORB
    seqa = syn.sequence(5, 640, 480, seed=41)          # one texture: frames 0..2 are stored, frame 4 comes back to it
    seqb = syn.sequence(3, 640, 480, seed=42)          # another place
    orb = ORB(nfeatures=500, max_size=(640, 480), max_batch=6)
    kps, desc, counts = orb.extract_batch(list(seqb) + list(seqa[:3]))
    cur = orb.extract_batch([seqa[4]])
    q = cur[1][0, :cur[2][0]]
    scores, best = bf.loop_score(q, desc, counts, 10, 30)
    want, wb = oracle.loop_score(q, desc, counts, 10, 30)
    assert np.array_equal(scores, want) and best == wb and best >= 3
    assert scores[3:].min() > 2 * max(scores[:3].max(), 1)
    orb.close()
