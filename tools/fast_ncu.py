"""Three batches of 64 frames through orbx_extract_batch_dev for ncu (tools/profile_step_r2.sh): `dense` = the Appendix-B stress
frames of the benchmark, `natural` = synthetic.natural_frame.  The third batch is the one to look at (K3's density hint set)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from monocular_slam_b200 import ORB
from monocular_slam_b200 import synthetic as syn

kind = sys.argv[1] if len(sys.argv) > 1 else "dense"
B, W, H = 64, 1920, 1080
gen = syn.natural_frame if kind == "natural" else syn.frame
seq = torch.from_numpy(syn.sequence(B, W, H, seed=100, generator=gen)).cuda()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
orb = ORB(nfeatures=2000, max_size=(W, H), max_batch=B)
orb.set_stream(stream.cuda_stream)
cap = orb.default_cap
kps = torch.empty((B, cap, 7), dtype=torch.float32, device="cuda")
desc = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
for _ in range(3):
    orb.extract_batch_dev(seq.data_ptr(), W * H, B, W, H, W, kps.data_ptr(), desc.data_ptr(), cap, cnt.data_ptr())
stream.synchronize()
orb.check_dev()
print(kind, "keypoints per frame", float(cnt.float().mean()))
orb.close()
