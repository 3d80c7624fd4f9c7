"""Scratch timing probe for gpurun sessions (not part of the product or the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monocular_slam_b200 import ORB, BFMatcher, popc_peak
from monocular_slam_b200 import synthetic as syn

def ev_time(fn, stream, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))

print(torch.cuda.get_device_name(0))
g, ms = popc_peak(0); print("popc peak %.1f Gpopc/s (%.2f ms) -> %.1f Gcmp/s" % (g, ms, g / 8))
s = torch.cuda.Stream()   # a non-default stream: handle 0 would mean "use the library's own stream"
torch.cuda.set_stream(s)
QUICK = len(sys.argv) > 1 and sys.argv[1] == "quick"
m = BFMatcher(); m.set_stream(s.cuda_stream)
for nq, nt in ([(2000, 2000), (65536, 125000)] if QUICK else [(2000, 2000), (8000, 8000), (2000, 200000), (65536, 125000), (262144, 125000), (1000000, 125000)]):
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda"); t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device="cuda")
    out = torch.empty((nq, 4), dtype=torch.int32, device="cuda")
    best, med = ev_time(lambda: m.knn2_dev(q.data_ptr(), nq, t.data_ptr(), nt, 0, out.data_ptr()), s, reps=2 if QUICK else (3 if nq * nt > 1e10 else 10))
    print("knn2 %8d x %8d: best %.3f ms med %.3f ms -> %.1f Gcmp/s" % (nq, nt, best, med, nq * nt / best / 1e6))
for (w, h, nf, B) in ([(1920, 1080, 2000, 8)] if QUICK else [(1920, 1080, 2000, 1), (1920, 1080, 2000, 8), (1920, 1080, 2000, 32), (1241, 376, 2000, 16), (3840, 2160, 8000, 4)]):
    seq = syn.sequence(min(B, 4), w, h, seed=1)
    frames = torch.from_numpy(np.stack([seq[i % len(seq)] for i in range(B)])).cuda()
    orb = ORB(nfeatures=nf, max_size=(w, h), max_batch=B); orb.set_stream(s.cuda_stream)
    cap = orb.default_cap
    kps = torch.empty((B, cap, 7), dtype=torch.float32, device="cuda"); desc = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda"); cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
    fn = lambda: orb.extract_batch_dev(frames.data_ptr(), w * h, B, w, h, w, kps.data_ptr(), desc.data_ptr(), cap, cnt.data_ptr())
    best, med = ev_time(fn, s, reps=3 if QUICK else 10)
    orb.check_dev()
    print("extract %dx%d nf=%d batch=%d: best %.3f ms med %.3f -> %.1f frames/s; counts %s" % (w, h, nf, B, best, med, B / best * 1e3, cnt[:3].tolist()))
    # host path
    hf = [seq[i % len(seq)] for i in range(B)]
    t0 = time.perf_counter(); 
    for _ in range(3): orb.extract_batch(hf)
    dt = (time.perf_counter() - t0) / 3
    print("   host-buffer batch: %.3f ms -> %.1f frames/s" % (dt * 1e3, B / dt))
    orb.close()
