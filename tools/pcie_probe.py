"""Scratch: pinned host -> device and device -> host copy bandwidth of the box (gpurun only)."""
import torch, time
n = 64 * 1920 * 1080
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10): fn()
        e1.record(s); e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("%s %d MB: %.3f ms -> %.1f GB/s" % (name, n // 1000000, ms, n / ms / 1e6))
