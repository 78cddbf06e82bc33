"""Quick HBM probes with stock torch kernels (context for the DBF kernel's achieved bandwidth)."""
import torch
torch.cuda.init()
n = 512 * 1024 * 1024 // 4
a = torch.randn(n, device="cuda")
b = torch.empty_like(a)
def timeit(f, reps=10):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    best = 1e9
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t = timeit(lambda: b.copy_(a)); print(f"copy 512MB: {2*a.numel()*4/t/1e6:.0f} GB/s (read+write)")
t = timeit(lambda: a.sum()); print(f"sum 512MB (read only): {a.numel()*4/t/1e6:.0f} GB/s")
t = timeit(lambda: b.fill_(1.0)); print(f"fill 512MB (write only): {a.numel()*4/t/1e6:.0f} GB/s")
c = a[: 67108864 // 4]
t = timeit(lambda: c.sum()); print(f"sum 67MB (read only, cold-ish): {c.numel()*4/t/1e6:.0f} GB/s")
