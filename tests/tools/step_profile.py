"""GPU: where does a training step spend its time?  (kernel-time table via torch.profiler + wall/GPU split)"""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
from oracle import gwnet_oracle as O
ge.build(); ge.load_package()
from graph_wavenet_b200 import engine as E, native as NV
from graph_wavenet_b200.metrics import StandardScaler
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
gen = torch.Generator().manual_seed(0)
sup = [s.to(dev) for s in O.synthetic_supports(207, 0.05, gen)]
torch.manual_seed(999)
tr = E.trainer(StandardScaler(54.0, 20.0), 2, 12, 207, 32, 0.3, 1e-3, 1e-4, dev, sup, True, True, None)
tr.model.precision = {"fp32": 0, "tf32": 1, "fp32x3": 3}[prec]
x, y = O.synthetic_batch(64, 207, 12, 2, gen)
x, y = x.to(dev), y.to(dev)
for _ in range(5):
    tr.train(x, y)
torch.cuda.synchronize()

def fwd_bwd():
    tr.model.train()
    out = tr.model(torch.nn.functional.pad(x, (1, 0, 0, 0)))
    out.sum().backward()

for name, fn in (("train step", lambda: tr.train(x, y)), ("fwd+bwd only", fwd_bwd)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(20):
        fn()
    e1.record(); t_host = (time.perf_counter() - t0) / 20
    torch.cuda.synchronize()
    print(f"[{prec}] {name}: host-enqueue {t_host*1e3:.2f} ms/iter, gpu {e0.elapsed_time(e1)/20:.2f} ms/iter", flush=True)

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.train(x, y)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=90))
