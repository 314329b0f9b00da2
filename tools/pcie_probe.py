import torch, time
for mb in (17, 136, 512):
    n = mb * 1000 * 1000
    d = torch.empty(n, dtype=torch.uint8, device='cuda'); h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"D2H {mb} MB pinned: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D {mb} MB pinned: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s")
