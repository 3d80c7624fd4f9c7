#!/usr/bin/env python
"""Device times of the bag-of-words kernels on the bench shape (64 frames x 2000 descriptors, k = 10, L = 6 vocabulary):
the descent alone (bowx_transform_features_dev), the whole transform (descent + k_bow_build), the scoring of one vector
against 10 000 stored ones.  Usage: python tools/bow_probe.py [nframes] [features per frame] [k] [L]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monocular_slam_b200 import Vocabulary  # noqa: E402
from monocular_slam_b200 import synthetic as syn  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    L = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    va = syn.vocabulary_large(77, k, L)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    leaves = np.flatnonzero(va["leaf"])
    desc = va["desc"][rng.choice(leaves, size=(B, n))].copy()
    desc[:, :, 3] ^= rng.integers(0, 4, (B, n), dtype=np.uint8)
    d_desc = torch.from_numpy(desc).to(dev)
    d_cnt = torch.full((B,), n, dtype=torch.int32, device=dev)
    voc = Vocabulary(va)
    stream = torch.cuda.Stream()
    voc.set_stream(stream.cuda_stream)
    w = torch.zeros((B, n), dtype=torch.int32, device=dev); wt = torch.zeros((B, n), dtype=torch.float64, device=dev); nid = torch.zeros_like(w)
    bw = torch.zeros_like(w); bv = torch.zeros_like(wt); bn = torch.zeros(B, dtype=torch.int32, device=dev)
    fnod = torch.zeros_like(w); foff = torch.zeros((B, n + 1), dtype=torch.int32, device=dev); ffe = torch.zeros_like(w); fn = torch.zeros_like(bn)

    def timed(fn_, reps=20):
        with torch.cuda.stream(stream):
            for _ in range(3):
                fn_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fn_()
            e1.record(stream)
            stream.synchronize()
        return e0.elapsed_time(e1) / reps

    t_desc = timed(lambda: voc.transform_features_dev(d_desc.data_ptr(), d_cnt.data_ptr(), B, n, 4, w.data_ptr(), wt.data_ptr(), nid.data_ptr()))
    t_bow = timed(lambda: voc.transform_batch_dev(d_desc.data_ptr(), d_cnt.data_ptr(), B, n, 4, bw.data_ptr(), bv.data_ptr(), bn.data_ptr()))
    t_all = timed(lambda: voc.transform_batch_dev(d_desc.data_ptr(), d_cnt.data_ptr(), B, n, 4, bw.data_ptr(), bv.data_ptr(), bn.data_ptr(),
                                                  fnod.data_ptr(), foff.data_ptr(), ffe.data_ptr(), fn.data_ptr()))
    print("%d frames x %d descriptors, k=%d L=%d (%d nodes): descent %.3f ms, + BowVector %.3f ms, + FeatureVector %.3f ms  (%.0f M features/s)"
          % (B, n, k, L, len(va["parent"]), t_desc, t_bow, t_all, B * n / t_all / 1e3))
    nb = bn.cpu().numpy()
    ndb = 10000
    per = int(nb.max())
    pw = bw[:, :per].repeat((ndb + B - 1) // B, 1)[:ndb].contiguous(); pv = bv[:, :per].repeat((ndb + B - 1) // B, 1)[:ndb].contiguous()
    pstart = torch.arange(ndb, dtype=torch.int64, device=dev) * per
    dcount = bn[(torch.arange(ndb, device=dev) % B)].contiguous()
    sc = torch.zeros(ndb, dtype=torch.float64, device=dev)
    t_sc = timed(lambda: voc.score_batch_dev(bw[0].data_ptr(), bv[0].data_ptr(), int(nb[0]), pstart.data_ptr(), dcount.data_ptr(), pw.data_ptr(),
                                             pv.data_ptr(), ndb, sc.data_ptr()))
    bytes_ = int(dcount.sum().item()) * 4
    print("score: 1 x %d stored vectors of %.0f words, every %dth one the query itself: %.3f ms = %.0f GB/s of stored word ids" % (ndb, nb.mean(), B, t_sc, bytes_ / t_sc / 1e6))
    # the same database without the copies of the query: entries that held frame 0 now hold frame 1
    pw2, pv2, dc2 = pw.clone(), pv.clone(), dcount.clone()
    pw2[0::B] = pw[1::B][:len(pw2[0::B])]; pv2[0::B] = pv[1::B][:len(pv2[0::B])]; dc2[0::B] = dcount[1::B][:len(dc2[0::B])]
    t_sc2 = timed(lambda: voc.score_batch_dev(bw[0].data_ptr(), bv[0].data_ptr(), int(nb[0]), pstart.data_ptr(), dc2.data_ptr(), pw2.data_ptr(),
                                              pv2.data_ptr(), ndb, sc.data_ptr()))
    print("score: the same without the copies of the query: %.3f ms = %.0f GB/s of stored word ids" % (t_sc2, int(dc2.sum().item()) * 4 / t_sc2 / 1e6))
    t_one = timed(lambda: voc.score_batch_dev(bw[0].data_ptr(), bv[0].data_ptr(), int(nb[0]), pstart.data_ptr(), dcount.data_ptr(), pw.data_ptr(),
                                              pv.data_ptr(), 1, sc.data_ptr()))
    t_none = timed(lambda: voc.score_batch_dev(bw[0].data_ptr(), bv[0].data_ptr(), int(nb[0]), pstart[1:].data_ptr(), dcount[1:].data_ptr(), pw.data_ptr(),
                                               pv.data_ptr(), 1, sc.data_ptr()))
    print("score: the query against itself alone (%d common words, added in order): %.3f ms; against one other vector: %.3f ms" % (nb[0], t_one, t_none))


if __name__ == "__main__":
    main()
