"""Generate tests/golden/diffg.npz from the REAL reference ``gwnet_diff_G`` (model.py:244-407).  Build container only.
The network draws fresh node embeddings inside forward (model.py:324-329); the script seeds the CPU generator right
before every forward and records the seed, so an implementation that draws in the same order reproduces the run."""
import os, sys
import numpy as np
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))
from make_golden import load_reference            # noqa: E402
from oracle import diffg_oracle as DO             # noqa: E402

ref_model, _ = load_reference()
N, B, T, S = 20, 4, 49, 2
kw = dict(dropout=0.0, supports_len=S + 1, gcn_bool=True, addaptadj=True, in_dim=2, out_dim=12, residual_channels=32,
          dilation_channels=32, skip_channels=64, end_channels=128, kernel_size=2, blocks=4, layers=2)
torch.manual_seed(5)
m = ref_model.gwnet_diff_G("cpu", N, **kw)
gen = torch.Generator().manual_seed(3)
x = torch.randn(B, 2, N, T, generator=gen)
sup = []
for _ in range(S):
    a = torch.rand(B, N, N, generator=gen) * (torch.rand(B, N, N, generator=gen) < 0.3).float() + torch.eye(N)
    sup.append(a / a.sum(dim=2, keepdim=True))
state0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
FWD_SEED = 77
m.train()
torch.manual_seed(FWD_SEED)
xr = x.clone().requires_grad_(True)
out = m(xr, [s.clone() for s in sup], None)
probe = torch.randn(out.shape, generator=gen)
(out * probe).sum().backward()
rec = {"x": x.numpy(), "probe": probe.numpy(), "out_train": out.detach().numpy(), "grad_input": xr.grad.numpy(),
       "fwd_seed": np.array(FWD_SEED), "cfg_N": np.array(N), "cfg_skip": np.array(64), "cfg_end": np.array(128)}
for i, s in enumerate(sup):
    rec[f"support.{i}"] = s.numpy()
for k, v in state0.items():
    rec["state0/" + k] = v.numpy()
for k, p in m.named_parameters():
    if p.grad is not None:
        rec["grad/" + k] = p.grad.numpy()
for k, v in m.state_dict().items():
    if "running" in k or "num_batches" in k:
        rec["buf1/" + k] = v.numpy()
m.eval()
torch.manual_seed(FWD_SEED + 1)
with torch.no_grad():
    rec["out_eval"] = m(x, [s.clone() for s in sup], None).numpy()
# the oracle restatement must reproduce the reference bit for bit from the same draws
st = {k: v.clone() for k, v in state0.items()}
torch.manual_seed(FWD_SEED)
nv = DO.draw_node_embeddings(B, N)
o2 = DO.forward(st, x, sup, nv, dropout=0.0, training=True)
print("oracle vs reference, train output max abs diff:", float((o2 - out.detach()).abs().max()))
assert torch.equal(o2, out.detach())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "diffg.npz"), **rec)
print("wrote diffg.npz:", {k: v.shape for k, v in rec.items() if not k.startswith(("state0/", "grad/", "buf1/"))},
      "grads:", sum(1 for k in rec if k.startswith("grad/")))
