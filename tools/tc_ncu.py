"""One launch pair for ncu: the tensor-core matcher on 75776 x 125000 (296 CTAs = 2 per SM, 977 train tiles each)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from monocular_slam_b200 import BFMatcher, _lib

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
m = BFMatcher()
m.set_stream(stream.cuda_stream)
nq, nt = 148 * 2 * 256, 125000
g = torch.Generator(device=dev)
g.manual_seed(1)
q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=dev, generator=g)
t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device=dev, generator=g)
o = torch.empty((nq, 4), dtype=torch.int32, device=dev)
m.set_kernel(_lib.KERNEL_INTEGER if len(sys.argv) > 1 and sys.argv[1] == "int" else _lib.KERNEL_TENSOR)
fn = m.knn2_dev
for _ in range(3):
    fn(q.data_ptr(), nq, t.data_ptr(), nt, 0, o.data_ptr())
stream.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
fn(q.data_ptr(), nq, t.data_ptr(), nt, 0, o.data_ptr())
e1.record(stream)
e1.synchronize()
print("%.3f ms  %.0f Gcmp/s" % (e0.elapsed_time(e1), nq * nt / e0.elapsed_time(e1) / 1e6))
m.close()
