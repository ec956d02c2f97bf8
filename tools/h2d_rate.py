import torch, time
n = 552_000_000 // 8
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print("H2D 552 MB pinned: %.2f ms  %.1f GB/s" % (e0.elapsed_time(e1), 0.552 / e0.elapsed_time(e1) * 1e3))
