"""Multi-process check of the sharded paths on real GPUs (run under torchrun, one rank per GPU):
* train-sharded matcher: peer-memory (fused scatter + flag-wait merge) and NCCL all-gather exchanges must both equal a
  single-device pass;
* frame-sharded loop-closure scoring: per-frame scores and the winning frame must equal a single-device pass.
Launched by tests/test_gpu_multi.py; also usable by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mp_sharded_check.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from monocular_slam_b200 import BFMatcher
from monocular_slam_b200.sharded import ShardedMatcher, shard_bounds


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    same = False      # one rank per GPU only: ranks that spin on each other's flags must not share a device
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    m = BFMatcher(device=local)
    m.set_stream(stream.cuda_stream)
    nq_max = 70000
    p2p = ShardedMatcher(m, p2p=True, nq_max=nq_max)
    nccl = None if same else ShardedMatcher(m, p2p=False)
    timings = {"p2p": float("nan"), "nccl": float("nan")}
    cases = [(2000, 200000), (1, 17), (257, 1000), (65536, 100003), (2000, 200000)]
    if same:
        cases = [(2000, 20000), (1, 17), (257, 1000), (20000, 3003)]      # time-sliced contexts: keep the spins short
    for case, (nq, nt) in enumerate(cases):
        g = torch.Generator(device=dev)
        g.manual_seed(1000 + case)                          # same data on every rank
        q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=dev, generator=g)
        t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device=dev, generator=g)
        t[nt // 3] = q[0]
        t[:: max(nt // 50, 1)] = t[0]                        # duplicates across shard borders
        b = shard_bounds(nt, world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        shard = t[lo:hi].contiguous()
        ref = torch.empty((nq, 4), dtype=torch.int32, device=dev)
        m.knn2_dev(q.data_ptr(), nq, t.data_ptr(), nt, 0, ref.data_ptr())
        for rep in range(3):
            a = p2p.knn2(q, shard, lo)
            c = nccl.knn2(q, shard, lo) if nccl is not None else ref
            stream.synchronize()
            assert torch.equal(a, ref), "rank %d case %d rep %d: peer-memory result differs from the single-device pass" % (rank, case, rep)
            assert torch.equal(c, ref), "rank %d case %d rep %d: all-gather result differs from the single-device pass" % (rank, case, rep)
        # device time of the two exchanges for the map-vs-frame shape (BASELINE config 4), max over ranks
        if (nq, nt) == (2000, 200000) and not same:
            for name, sm in (("p2p", p2p), ("nccl", nccl)):
                for _ in range(5):
                    sm.knn2(q, shard, lo)
                stream.synchronize()
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(50):
                    sm.knn2(q, shard, lo)
                e1.record(stream)
                e1.synchronize()
                mine = e0.elapsed_time(e1) / 50
                ms = torch.tensor([mine], dtype=torch.float64, device=dev)
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                timings[name] = float(ms.item())
                print("rank %d %s %.3f us" % (rank, name, mine * 1e3), file=sys.stderr, flush=True)
    # frame-sharded loop-closure scoring: 37 stored frames (uneven blocks) of up to 600 descriptors
    from monocular_slam_b200.sharded import ShardedLoopScorer
    g = torch.Generator(device=dev)
    g.manual_seed(77)
    nf, cap, nq = 37, 600, 500
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=dev, generator=g)
    frames = torch.randint(0, 256, (nf, cap, 32), dtype=torch.uint8, device=dev, generator=g)
    counts = torch.randint(0, cap + 1, (nf,), dtype=torch.int32, device=dev, generator=g)
    for f in (9, 30):
        frames[f, :nq] = q
        frames[f, :nq, 0] ^= 3
        counts[f] = cap
    want = torch.empty(nf, dtype=torch.int32, device=dev)
    wbest = torch.empty(2, dtype=torch.int32, device=dev)
    m.loop_score_dev(q.data_ptr(), nq, frames.data_ptr(), counts.data_ptr(), nf, cap, 10, 40, want.data_ptr(), wbest.data_ptr())
    b = shard_bounds(nf, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    sl = ShardedLoopScorer(m, n=10, thr=40)
    scores, best = sl.score(q, frames[lo:hi].contiguous(), counts[lo:hi].contiguous(), nf)
    stream.synchronize()
    assert torch.equal(scores, want) and torch.equal(best, wbest) and int(best[0]) == 9, "rank %d: frame-sharded loop scores differ" % rank
    # entry-sharded bag-of-words scoring: 41 stored vectors (uneven blocks) built on this device from seeded descriptors
    from monocular_slam_b200 import Vocabulary
    from monocular_slam_b200 import synthetic as syn
    from monocular_slam_b200.sharded import ShardedBowDatabase
    va = syn.vocabulary(13, k=9, L=3)
    voc = Vocabulary(va, device=dev.index)
    voc.set_stream(stream.cuda_stream)
    nent, bcap = 41, 500
    bdesc = np.stack([syn.vocabulary_features(200 + e, va, bcap, pool=120) for e in range(nent)])
    bcnt = torch.from_numpy(np.array([bcap if e % 5 else 37 * (e % 3) for e in range(nent)], np.int32)).to(dev)
    d_bdesc = torch.from_numpy(bdesc).to(dev)
    bw = torch.zeros((nent, bcap), dtype=torch.int32, device=dev); bv = torch.zeros((nent, bcap), dtype=torch.float64, device=dev)
    bn = torch.zeros(nent, dtype=torch.int32, device=dev)
    voc.transform_batch_dev(d_bdesc.data_ptr(), bcnt.data_ptr(), nent, bcap, 0, bw.data_ptr(), bv.data_ptr(), bn.data_ptr())
    bstart = torch.arange(nent, dtype=torch.int64, device=dev) * bcap
    bwant = torch.empty(nent, dtype=torch.float64, device=dev)
    qe = 23
    nq_ = int(bn[qe])
    voc.score_batch_dev(bw[qe].data_ptr(), bv[qe].data_ptr(), nq_, bstart.data_ptr(), bn.data_ptr(), bw.data_ptr(), bv.data_ptr(), nent, bwant.data_ptr())
    b = shard_bounds(nent, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    sdb = ShardedBowDatabase(voc)
    bscores, bbest = sdb.score(bw[qe, :nq_].contiguous(), bv[qe, :nq_].contiguous(), bstart[lo:hi].contiguous(), bn[lo:hi].contiguous(), bw, bv, nent)
    stream.synchronize()
    assert torch.equal(bscores, bwant) and bbest == qe and abs(float(bwant[qe]) - 1.0) < 1e-12, "rank %d: entry-sharded bag-of-words scores differ" % rank
    voc.close()
    p2p.close()
    m.close()
    dist.barrier()
    if rank == 0:
        print("MP_SHARDED_OK world=%d  2000x200000: peer-memory %.1f us, nccl all-gather %.1f us per call" % (
            world, timings["p2p"] * 1e3, timings["nccl"] * 1e3), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
