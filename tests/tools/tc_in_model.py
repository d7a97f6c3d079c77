import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
from oracle import gwnet_oracle as O
ge.build(); ge.load_package()
from graph_wavenet_b200 import model as M
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)
sup = [s.to(dev) for s in O.synthetic_supports(207, 0.05, gen)]
torch.manual_seed(999)
m = M.gwnet(dev, 207, 0.3, supports=sup).to(dev)
m.precision = 1
x, _ = O.synthetic_batch(64, 207, 12, 2, gen)
x = torch.nn.functional.pad(x, (1, 0, 0, 0)).to(dev)
m.train()
for _ in range(3):
    m(x).sum().backward()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m(x).sum().backward()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if "gwn" in e.name or "Memset" in e.name]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print(f"{(e.time_range.start - t0):9.1f} us  dur {e.time_range.end - e.time_range.start:8.1f} us  {e.name[:110]}")
