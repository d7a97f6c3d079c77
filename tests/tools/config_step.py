"""Minimal driver for ncu launch lists at another configuration: 3 un-captured trainer.train steps.
usage: python tests/tools/config_step.py <N> <B> <T> [tier]"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
from oracle import gwnet_oracle as O
ge.build(); ge.load_package()
from graph_wavenet_b200 import engine as E, native as NV
from graph_wavenet_b200.metrics import StandardScaler
N, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
tier = sys.argv[4] if len(sys.argv) > 4 else "fp32x3"
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)
sup = [s.to(dev) for s in O.synthetic_supports(N, 0.05, gen)]
x, y = O.synthetic_batch(B, N, T, 2, gen)
x, y = x.to(dev), y.to(dev)
torch.manual_seed(999)
tr = E.trainer(StandardScaler(54.0, 20.0), 2, T, N, 32, 0.3, 1e-3, 1e-4, dev, sup, True, True, None)
tr.model.precision = {"tf32": NV.PREC_TF32, "fp32x3": NV.PREC_FP32X3}[tier]
tr.use_graph = False
for _ in range(3):
    tr.train(x, y)
torch.cuda.synchronize()
print("done")
