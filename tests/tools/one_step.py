"""Minimal driver for ncu: 2 forward+backward passes of the METR-LA model (first = warm-up)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
from oracle import gwnet_oracle as O
ge.load_package()
from graph_wavenet_b200 import model as M
dev = torch.device("cuda:0")
prec = {"fp32": 0, "tf32": 1, "fp32x3": 3}[sys.argv[1] if len(sys.argv) > 1 else "tf32"]
gen = torch.Generator().manual_seed(0)
sup = [s.to(dev) for s in O.synthetic_supports(207, 0.05, gen)]
torch.manual_seed(999)
m = M.gwnet(dev, 207, 0.3, supports=sup).to(dev)
m.precision = prec
x, _ = O.synthetic_batch(64, 207, 12, 2, gen)
x = torch.nn.functional.pad(x, (1, 0, 0, 0)).to(dev)
m.train()
for _ in range(2):
    m(x).sum().backward()
torch.cuda.synchronize()
print("done")
