"""GPU: measured TF32 tensor-core peak of this box (cuBLAS through torch.matmul with allow_tf32, 8192^3, best of 10 and a
4 s sustained loop) -- the denominator for the tf32-tier tensor-pipe fractions (MEASURED_PEAKS.json only has bf16)."""
import time, json, torch
torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda:0")
n = 8192
a = torch.randn(n, n, device=dev); b = torch.randn(n, n, device=dev)
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a @ b; e1.record(); e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
t0 = time.perf_counter(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); k = 0
while time.perf_counter() - t0 < 4.0:
    for _ in range(10):
        a @ b
    k += 10
    torch.cuda.synchronize()
e1.record(); e1.synchronize()
sus = e0.elapsed_time(e1) / k
print(json.dumps({"tf32_tflops_burst": round(2 * n ** 3 / best / 1e9, 1), "tf32_tflops_sustained": round(2 * n ** 3 / sus / 1e9, 1),
                  "how": "torch.matmul fp32 inputs, allow_tf32=True, 8192^3"}))
