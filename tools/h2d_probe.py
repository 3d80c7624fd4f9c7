import torch, time
s = torch.cuda.Stream()
for mb in (1, 5, 19, 40, 133):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(s):
        for _ in range(3): d.copy_(h, non_blocking=True)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10): d.copy_(h, non_blocking=True)
        e1.record(s); s.synchronize()
    ms = e0.elapsed_time(e1) / 10
    # again with the host buffer rewritten before every copy
    with torch.cuda.stream(s):
        tot = 0
        for _ in range(10):
            h.add_(1)
            e0.record(s); d.copy_(h, non_blocking=True); e1.record(s); s.synchronize(); tot += e0.elapsed_time(e1)
    print("%4d MB: %.3f ms = %.1f GB/s; rewritten before each copy: %.3f ms = %.1f GB/s" % (mb, ms, n / ms / 1e6, tot / 10, n / (tot / 10) / 1e6))
