"""Probe of the tensor-core matcher (hamx_knn2_tc_dev) against the integer-pipe kernel (hamx_knn2_dev): parity on a few
shapes, then timing.  Run under `timeout` on a GPU box:  timeout 120 python tools/tc_probe.py [big]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from monocular_slam_b200 import BFMatcher, _lib

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
m = BFMatcher()
m.set_stream(stream.cuda_stream)
mt = BFMatcher()
mt.set_stream(stream.cuda_stream)
m.set_kernel(_lib.KERNEL_INTEGER)
mt.set_kernel(_lib.KERNEL_TENSOR)
m.knn2_tc_dev = mt.knn2_dev


def run(nq, nt, seed=0, dup=True):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=dev, generator=g)
    t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device=dev, generator=g)
    if dup and nt > 300 and nq > 4:
        t[100] = q[0]; t[200] = q[0]; t[17] = q[1]; t[nt - 1] = q[2]; t[0] = q[3]
    a = torch.full((nq, 4), -7, dtype=torch.int32, device=dev)
    b = torch.full((nq, 4), -9, dtype=torch.int32, device=dev)
    m.knn2_dev(q.data_ptr(), nq, t.data_ptr(), nt, 5, a.data_ptr())
    m.knn2_tc_dev(q.data_ptr(), nq, t.data_ptr(), nt, 5, b.data_ptr())
    stream.synchronize()
    bad = (a != b).any(dim=1).nonzero().flatten()
    print("nq=%d nt=%d: %d mismatching rows" % (nq, nt, bad.numel()), flush=True)
    if bad.numel():
        for i in bad[:5].tolist():
            print("  row", i, "ref", a[i].tolist(), "tc", b[i].tolist())
    return bad.numel() == 0


ok = True
for nq, nt in [(256, 128), (256, 384), (1, 1), (300, 1000), (1000, 130), (2000, 2000), (5000, 20000), (777, 12345), (70000, 3000)]:
    ok &= run(nq, nt, seed=nq + nt)
print("PARITY", "OK" if ok else "FAILED", flush=True)
if ok and len(sys.argv) > 1:
    nq, nt = 1 << 20, 125000
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=dev, generator=g)
    t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device=dev, generator=g)
    a = torch.empty((nq, 4), dtype=torch.int32, device=dev)
    b = torch.empty((nq, 4), dtype=torch.int32, device=dev)
    for name, fn, out in (("int-pipe", m.knn2_dev, a), ("tensor-core", m.knn2_tc_dev, b)):
        fn(q.data_ptr(), nq, t.data_ptr(), nt, 0, out.data_ptr())
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            fn(q.data_ptr(), nq, t.data_ptr(), nt, 0, out.data_ptr())
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("%s: %.2f ms  %.0f Gcmp/s" % (name, ms, nq * nt / ms / 1e6), flush=True)
    print("big parity:", bool(torch.equal(a, b)), flush=True)
    for nq2, nt2 in ((2000, 200000), (2000, 25000)):
        q2, t2 = q[:nq2].contiguous(), t[:nt2].contiguous() if nt2 <= nt else torch.randint(0, 256, (nt2, 32), dtype=torch.uint8, device=dev, generator=g)
        o = torch.empty((nq2, 4), dtype=torch.int32, device=dev)
        for name, fn in (("int-pipe", m.knn2_dev), ("tensor-core", m.knn2_tc_dev)):
            for _ in range(3):
                fn(q2.data_ptr(), nq2, t2.data_ptr(), nt2, 0, o.data_ptr())
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(20):
                fn(q2.data_ptr(), nq2, t2.data_ptr(), nt2, 0, o.data_ptr())
            e1.record(stream)
            e1.synchronize()
            print("%d x %d %s: %.1f us" % (nq2, nt2, name, e0.elapsed_time(e1) / 20 * 1e3), flush=True)
m.close()
