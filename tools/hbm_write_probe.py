import torch, time
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for mb in (140, 512, 2048):
    n = mb*1024*1024//4
    a = torch.empty(n, dtype=torch.int32, device='cuda'); b = torch.empty(n, dtype=torch.int32, device='cuda')
    ms = t(lambda: a.fill_(7)); print(f'fill  {mb} MB: {ms*1e3:.1f} us  {mb*1.048576/ms:.0f} GB/s write-only')
    ms = t(lambda: b.copy_(a)); print(f'copy  {mb} MB: {ms*1e3:.1f} us  {2*mb*1.048576/ms:.0f} GB/s r+w')
    ms = t(lambda: a.sum()); print(f'sum   {mb} MB: {ms*1e3:.1f} us  {mb*1.048576/ms:.0f} GB/s read-only')
