"""GPU: per-operator table and launch summary of gwnet_diff_G (per-sample graphs) at the fork's default shape
(train.py: num_nodes 80, batch 32, seq_length 48, nhid 32), forward + backward through the autograd node."""
import ctypes, json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build(); ge.load_package()
from graph_wavenet_b200 import model as M, native as NV
lib = NV.get_lib()
dev = torch.device("cuda:0")
N, B, T = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (80, 32, 48)))
tier = sys.argv[4] if len(sys.argv) > 4 else "fp32x3"
torch.manual_seed(5)
m = M.gwnet_diff_G(dev, N, dropout=0.3, supports_len=3, out_dim=12).to(dev)
m.precision = {"tf32": NV.PREC_TF32, "fp32x3": NV.PREC_FP32X3, "fp32": NV.PREC_FP32}[tier]
gen = torch.Generator().manual_seed(3)
x = torch.randn(B, 2, N, T + 1, generator=gen).to(dev)
sup = []
for _ in range(2):
    a = torch.rand(B, N, N, generator=gen) * (torch.rand(B, N, N, generator=gen) < 0.3).float() + torch.eye(N)
    sup.append((a / a.sum(dim=2, keepdim=True)).to(dev))
m.train()


def step():
    m.zero_grad(set_to_none=True)
    out = m(x, sup, None)
    out.sum().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"gwnet_diff_G": [N, B, T], "tier": tier, "ms_fwd_bwd": e0.elapsed_time(e1) / 10}))
lib.check(lib.dll.gwn_profile_begin())
for _ in range(3):
    step()
buf = ctypes.create_string_buffer(1 << 16)
lib.check(lib.dll.gwn_profile_end(buf, len(buf)))
ops = json.loads(buf.value.decode())
for o in sorted(ops, key=lambda o: -o["ms"]):
    print(f"{o['op']:22s} {o['ms'] / 3 * 1e3:9.1f} us  calls/step {o['calls'] / 3:5.1f}")
