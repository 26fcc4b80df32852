import torch, time
for mb in (10, 78, 236):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for direction in ("h2d", "d2h"):
        ts = []
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if direction == "h2d": d.copy_(h, non_blocking=True)
            else: h.copy_(d, non_blocking=True)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print("%s %4d MiB  %.3f ms  %.1f GB/s" % (direction, mb, t, (mb << 20) / t / 1e6))
