"""Pipelined host path with and without the outlier filter: wall-clock ms per 64-frame step (tools; GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from monocular_slam_b200 import ORB, BFMatcher, DMATCH_DTYPE, KEYPOINT_DTYPE, FundamentalFilter
from monocular_slam_b200 import synthetic as syn

B, W, H = 64, 1920, 1080
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
seq = torch.from_numpy(syn.sequence(B, W, H, seed=100)).pin_memory()
frames = [seq[i].numpy() for i in range(B)]
orb = ORB(nfeatures=2000, max_size=(W, H), max_batch=B)
m, fm = BFMatcher(), FundamentalFilter()
cap, depth = orb.default_cap, orb.pipeline_depth()


def pin(shape, dt):
    return torch.zeros(shape, dtype=dt).pin_memory().numpy()


def outs(back):
    o = (pin((B, cap, 7), torch.float32).view(KEYPOINT_DTYPE).reshape(B, cap), pin((B, cap, 32), torch.uint8), np.zeros(B, np.int32))
    if back:
        return o + (pin((B, back, cap, 4), torch.int32).view(DMATCH_DTYPE).reshape(B, back, cap), np.zeros((B, back), np.int64),
                    pin((B, back, cap), torch.uint8), pin((B, back, 3, 3), torch.float64), np.zeros((B, back), np.int32))
    return o + (pin((B, cap, 4), torch.int32).view(DMATCH_DTYPE).reshape(B, cap), np.zeros(B, np.int64), pin((B, cap), torch.uint8),
                pin((B, 3, 3), torch.float64), np.zeros(B, np.int32))


for name, f, back in (("plain", None, 0), ("filtered", fm, 0), ("back5", None, 5), ("back5+filter", fm, 5), ("plain", None, 0)):
    bufs = [outs(back) for _ in range(depth)]
    orb.reset_sequence()
    k = 0

    def step():
        global k
        if orb.batches_in_flight() == depth:
            orb.wait_batch()
        orb.submit_batch(frames, m, 0.75, bufs[k % depth], fundamental=f, back=back)
        k += 1
    for _ in range(4):
        step()
    while orb.batches_in_flight():
        orb.wait_batch()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    while orb.batches_in_flight():
        orb.wait_batch()
    print("%-14s %.3f ms per step" % (name, (time.perf_counter() - t0) * 1e3 / steps), flush=True)
