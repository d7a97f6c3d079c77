"""GPU: per-operator table (bench.roofline_leg) of a full trainer.train step at another BASELINE configuration.
usage: python tests/tools/config_ops.py <N> <B> <T> [tier]      e.g. 200 32 48 fp32x3   (CRASH shape, the fork's seq_length 48)"""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import bench
import __graft_entry__ as ge
from oracle import gwnet_oracle as O
ge.build(); ge.load_package()
from graph_wavenet_b200 import engine as E, native as NV
from graph_wavenet_b200.metrics import StandardScaler
N, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
tier = sys.argv[4] if len(sys.argv) > 4 else "fp32x3"
dev = torch.device("cuda:0")
lib = NV.get_lib()
gen = torch.Generator().manual_seed(0)
sup = [s.to(dev) for s in O.synthetic_supports(N, 0.05, gen)]
x, y = O.synthetic_batch(B, N, T, 2, gen)
x, y = x.to(dev), y.to(dev)
torch.manual_seed(999)
tr = E.trainer(StandardScaler(54.0, 20.0), 2, T, N, 32, 0.3, 1e-3, 1e-4, dev, sup, True, True, None)
tr.model.precision = {"tf32": NV.PREC_TF32, "fp32x3": NV.PREC_FP32X3}[tier]
bench.PRECISION_FOR_NOTE[0] = tier
for _ in range(3):
    tr.train(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    tr.train(x, y)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"N": N, "B": B, "T": T, "tier": tier, "ms_per_step_graph": e0.elapsed_time(e1) / 10}))
tr.use_graph = False
roof, table = bench.roofline_leg(lib, lambda i: tr.train(x, y), dev, steps=3)
for o in table:
    print(f"{o['op']:22s} {o['ms_per_step']*1e3:9.1f} us  share {o['share']:.3f} hbm {o['hbm_frac']:.2f} tens {o['tensor_frac']:.2f} {o['bound']}")
