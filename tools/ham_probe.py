"""Scratch: matcher throughput at the BASELINE sizes (gpurun only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monocular_slam_b200 import BFMatcher, popc_peak
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
g, ms = popc_peak(0); print("popc peak %.1f Gpopc/s -> %.1f Gcmp/s" % (g, g / 8))
m = BFMatcher(); m.set_stream(s.cuda_stream)
for nq, nt in [(2000, 2000), (8000, 8000), (2000, 200000), (262144, 125000)]:
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda"); t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device="cuda")
    out = torch.empty((nq, 4), dtype=torch.int32, device="cuda")
    fn = lambda: m.knn2_dev(q.data_ptr(), nq, t.data_ptr(), nt, 0, out.data_ptr())
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); fn(); e1.record(s); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    print("knn2 %8d x %8d: best %.3f ms -> %.1f Gcmp/s" % (nq, nt, min(ts), nq * nt / min(ts) / 1e6))
# 64 batched 2000x2000 consecutive-frame problems (the headline step's matching)
B, cap = 64, 2500
desc = torch.randint(0, 256, (B, cap, 32), dtype=torch.uint8, device="cuda"); cnt = torch.full((B,), 2000, dtype=torch.int32, device="cuda")
good = torch.empty((B, cap, 4), dtype=torch.int32, device="cuda"); ngood = torch.zeros(B, dtype=torch.int64, device="cuda")
prev = torch.randint(0, 256, (cap, 32), dtype=torch.uint8, device="cuda"); prevn = torch.full((1,), 2000, dtype=torch.int32, device="cuda")
fn = lambda: m.match_consecutive_dev(desc.data_ptr(), cnt.data_ptr(), B, cap, prev.data_ptr(), prevn.data_ptr(), 0.75, good.data_ptr(), ngood.data_ptr())
for _ in range(2): fn()
torch.cuda.synchronize(); ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); fn(); e1.record(s); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print("64 x (2000 x 2000) batched: best %.3f ms -> %.1f Gcmp/s" % (min(ts), 64 * 4e6 / min(ts) / 1e6))
