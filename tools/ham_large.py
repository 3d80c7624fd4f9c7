"""One large train-sharded-shape matching call (262144 x 125000) for an ncu capture of k_hamming_knn2 (gpurun only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monocular_slam_b200 import BFMatcher
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
m = BFMatcher(); m.set_stream(s.cuda_stream)
nq, nt = 262144, 125000
q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda"); t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device="cuda")
out = torch.empty((nq, 4), dtype=torch.int32, device="cuda")
for _ in range(2):
    m.knn2_dev(q.data_ptr(), nq, t.data_ptr(), nt, 0, out.data_ptr())
s.synchronize()
print("done")
