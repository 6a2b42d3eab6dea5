import torch
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    best=1e9
    for _ in range(n):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
a=torch.empty(1<<30, dtype=torch.float32, device='cuda'); b=torch.empty_like(a)
ms=t(lambda: a.zero_()); print(f"memset 4GiB: {ms:.3f} ms -> {a.numel()*4/ms/1e6:.0f} GB/s")
ms=t(lambda: b.copy_(a)); print(f"copy 4GiB: {ms:.3f} ms -> {2*a.numel()*4/ms/1e6:.0f} GB/s (r+w)")
ms=t(lambda: a.sum()); print(f"read-reduce 4GiB: {ms:.3f} ms -> {a.numel()*4/ms/1e6:.0f} GB/s")
x=a[:10240*65536].view(10240,65536)
ms=t(lambda: x.fill_(1.0)); print(f"fill 2.7GB: {ms:.3f} ms -> {x.numel()*4/ms/1e6:.0f} GB/s")
