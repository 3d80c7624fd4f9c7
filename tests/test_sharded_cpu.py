"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU paths (ranges, global offsets, the single
all-gather, merge order).  The per-shard top-2 and the merge are CUDA kernels in the product; here the oracle and a
numpy lexicographic merge stand in for them so the plumbing can be checked without a GPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from monocular_slam_b200 import synthetic as syn
from monocular_slam_b200.sharded import ShardedMatcher, frame_block, shard_bounds


def test_shard_bounds():
    assert shard_bounds(10, 3).tolist() == [0, 4, 7, 10]
    assert shard_bounds(200000, 8).tolist() == [25000 * i for i in range(9)]
    assert shard_bounds(2, 4).tolist() == [0, 1, 2, 2, 2]
    covered = []
    for r in range(4):
        lo, hi, prev = frame_block(257, r, 4)
        covered += list(range(lo, hi))
        assert prev == (lo - 1 if lo else None)
    assert covered == list(range(257))


def _local_top2(q, t, offset):
    idx, d = oracle.knn2(q.numpy(), t.numpy())
    idx = np.where(idx >= 0, idx + offset, -1)
    return torch.from_numpy(np.stack([d[:, 0], idx[:, 0], d[:, 1], idx[:, 1]], 1).astype(np.int32))


def _merge(parts):
    p = parts.numpy().astype(np.int64)                      # [world, nq, 4]
    keys = np.concatenate([np.where(p[..., 1] >= 0, (p[..., 0] << 32) | p[..., 1], np.iinfo(np.int64).max),
                           np.where(p[..., 3] >= 0, (p[..., 2] << 32) | p[..., 3], np.iinfo(np.int64).max)], 0)
    keys.sort(axis=0)
    out = np.full((p.shape[1], 4), -1, np.int32)
    for j in range(2):
        k = keys[j]
        ok = k != np.iinfo(np.int64).max
        out[ok, 2 * j] = (k[ok] >> 32).astype(np.int32)
        out[ok, 2 * j + 1] = (k[ok] & 0xFFFFFFFF).astype(np.int32)
    return torch.from_numpy(out)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = syn.descriptors(31, 4001)
    t[3000:3005] = t[10:15]                     # duplicates on both sides of the shard boundary: tie rule across shards
    q = syn.planted_queries(32, t, 300)
    sm = ShardedMatcher(matcher=None, local_top2=_local_top2, merge=_merge)
    r = sm.knn2_from_full(torch.from_numpy(q), torch.from_numpy(t)).numpy()
    oi, od = oracle.knn2(q, t)
    ok = np.array_equal(r[:, [1, 3]], oi) and np.array_equal(r[:, [0, 2]], od)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_train_sharded_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    oracle.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_sharded_sequence_blocks_cpu():
    """ShardedSequence block logic with an injected extract/match (no GPU): every frame is reported exactly once, by the
    rank that owns it, and each block's first frame is matched against the previous rank's last frame (lead-in frame)."""
    from monocular_slam_b200.sharded import ShardedSequence
    nframes, world = 23, 4
    frames = list(range(nframes))        # a "frame" is its index; the stand-in's "match" records the frame it was matched to
    calls = []

    def fake(batch, first_is_lead_in):
        out = []
        for i, f in enumerate(batch):
            prev = None if (i == 0 and first_is_lead_in) else f - 1
            out.append(("kp%d" % f, "desc%d" % f, prev))
        calls.append((list(batch), first_is_lead_in))
        return out

    seen = {}
    for r in range(world):
        ss = ShardedSequence(rank=r, world=world, batch=4, extract_match=fake)
        lo, hi, prev = ss.block(nframes)
        got = list(ss.run(frames))
        assert [g[0] for g in got] == list(range(lo, hi))
        for f, k, d, m in got:
            assert f not in seen
            seen[f] = m
            assert k == "kp%d" % f and d == "desc%d" % f
    assert sorted(seen) == list(range(nframes))
    assert seen[0] is None and all(seen[f] == f - 1 for f in range(1, nframes))


def _loop_worker(rank, world, port, ret):
    """Frame-sharded loop-closure scoring over gloo with the oracle standing in for the scoring kernel."""
    from monocular_slam_b200.sharded import ShardedLoopScorer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = np.random.default_rng(5)
    nf, cap, nq, n, thr = 7, 40, 30, 10, 110             # 7 frames over 2 ranks: uneven blocks (4 + 3)
    q = r.integers(0, 256, (nq, 32), dtype=np.uint8)
    frames = r.integers(0, 256, (nf, cap, 32), dtype=np.uint8)
    counts = np.array([40, 12, 0, 40, 33, 40, 5], np.int32)
    frames[5, :15] = q[:15]
    frames[3, :15] = q[:15]

    def local(qt, ft, ct):
        return torch.from_numpy(oracle.loop_score(qt.numpy(), ft.numpy(), ct.numpy(), n, thr)[0].astype(np.int32))

    def best(sc):
        s = sc.numpy()
        return torch.tensor([int(np.argmax(s)) if s.max() > 0 else -1, int(s.max())], dtype=torch.int32)
    b = shard_bounds(nf, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    sl = ShardedLoopScorer(matcher=None, n=n, thr=thr, local_scores=local, best=best)
    scores, bst = sl.score(torch.from_numpy(q), torch.from_numpy(frames[lo:hi]), torch.from_numpy(counts[lo:hi]), nf)
    want, wb = oracle.loop_score(q, frames, counts, n, thr)
    ret[rank] = bool(np.array_equal(scores.numpy(), want) and int(bst[0]) == wb == 3)
    dist.destroy_process_group()


def test_frame_sharded_loop_scoring_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    oracle.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_loop_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def _bow_worker(rank, world, port, ret):
    """Entry-sharded bag-of-words scoring over gloo with the oracle standing in for the scoring kernel."""
    from monocular_slam_b200.sharded import ShardedBowDatabase
    from monocular_slam_b200 import synthetic as syn
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    va = syn.vocabulary(3, k=6, L=3)
    ov = oracle.BowVocabulary(va)
    bows = [ov.transform(syn.vocabulary_features(20 + e, va, n, pool=60)) for e, n in enumerate([200, 0, 150, 1, 90, 200, 33])]   # 7 entries: 4 + 3
    query = ov.transform(syn.vocabulary_features(24, va, 90, pool=60))          # entry 4's descriptors: score 1
    nent = len(bows)
    b = shard_bounds(nent, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    count = np.array([len(bw[0]) for bw in bows[lo:hi]], np.int32)
    start = np.concatenate([[0], np.cumsum(count[:-1])]).astype(np.int64)
    words = np.concatenate([bw[0] for bw in bows[lo:hi]]).astype(np.int64)
    vals = np.concatenate([bw[1] for bw in bows[lo:hi]])

    def local(qw, qv, st, ct, w, v):
        out = [ov.score((qw.numpy().astype(np.uint32), qv.numpy()), (w.numpy()[s:s + c].astype(np.uint32), v.numpy()[s:s + c])) for s, c in zip(st.tolist(), ct.tolist())]
        return torch.tensor(out, dtype=torch.float64)
    db = ShardedBowDatabase(None, local_scores=local)
    scores, best = db.score(torch.from_numpy(query[0].astype(np.int64)), torch.from_numpy(query[1]), torch.from_numpy(start), torch.from_numpy(count),
                            torch.from_numpy(words), torch.from_numpy(vals), nent)
    want = np.array([ov.score(query, bw) for bw in bows])
    ret[rank] = bool(np.array_equal(scores.numpy(), want) and best == 4 and abs(want[4] - 1.0) < 1e-12)
    dist.destroy_process_group()


def test_entry_sharded_bow_scoring_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    oracle.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bow_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
