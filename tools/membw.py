"""HBM micro-benchmark: pure write / pure read / copy bandwidth (is a write-heavy kernel bound below the copy roofline?)"""
import torch
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N = 1 << 30
a = torch.empty(N, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
ms = t(lambda: a.fill_(1.0)); print(f"write-only  fill_: {2*N/ms/1e6:.0f} GB/s")
ms = t(lambda: a.zero_()); print(f"write-only  memset: {2*N/ms/1e6:.0f} GB/s")
ms = t(lambda: b.copy_(a)); print(f"copy 1R+1W: {4*N/ms/1e6:.0f} GB/s")
ms = t(lambda: a.sum()); print(f"read-only sum: {2*N/ms/1e6:.0f} GB/s")
c = torch.empty(N // 4, dtype=torch.bfloat16, device="cuda")
ms = t(lambda: torch.cat([c, c, c, c], out=a)); print(f"1R(L2-ish 0.5GB)+4W cat: {(2*N + 2*N)/ms/1e6:.0f} GB/s total, writes {2*N/ms/1e6:.0f} GB/s")
