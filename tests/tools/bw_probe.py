"""Read-only vs read+write HBM bandwidth with library kernels (a yardstick for the roofline denominators)."""
import torch
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda").normal_()
y = torch.empty_like(x)
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
gb = x.numel() * 4 / 1e9
print("sum (read only)    %.0f GB/s" % (gb / t(lambda: x.sum()) * 1e3))
print("max (read only)    %.0f GB/s" % (gb / t(lambda: x.max()) * 1e3))
print("copy (read+write)  %.0f GB/s" % (2 * gb / t(lambda: y.copy_(x)) * 1e3))
print("fill (write only)  %.0f GB/s" % (gb / t(lambda: y.fill_(1.0)) * 1e3))
