"""Scratch (torchrun): per-rank timing of the 1M x 125k sharded matching step: kernel alone, peer-memory exchange, NCCL exchange."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from monocular_slam_b200 import BFMatcher
from monocular_slam_b200.sharded import ShardedMatcher
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
m = BFMatcher(device=local); m.set_stream(s.cuda_stream)
nq, nt = 1 << 20, 125000
q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=dev); t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device=dev)
out = torch.empty((nq, 4), dtype=torch.int32, device=dev)
p2p = ShardedMatcher(m, p2p=True, nq_max=nq); nccl = ShardedMatcher(m, p2p=False)
def timeit(fn, n=3):
    fn(); s.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(n): fn()
    e1.record(s); e1.synchronize()
    return e0.elapsed_time(e1) / n
a = timeit(lambda: m.knn2_dev(q.data_ptr(), nq, t.data_ptr(), nt, rank * nt, out.data_ptr()))
b = timeit(lambda: p2p.knn2(q, t, rank * nt))
c = timeit(lambda: nccl.knn2(q, t, rank * nt))
import pynvml; pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(local)
print("rank %d: kernel %.1f ms, peer-memory %.1f ms, nccl %.1f ms; clock %d MHz, power %.0f W" % (rank, a, b, c, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3), flush=True)
p2p.close(); m.close(); dist.barrier(); dist.destroy_process_group()
