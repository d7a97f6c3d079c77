"""GPU diagnostic: per-parameter gradient error of the CUDA path vs the fp64 oracle for a few configs."""
import sys, os
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
from oracle import gwnet_oracle as O

ge.build(); ge.load_package()
from graph_wavenet_b200 import model as M
dev = torch.device("cuda:0")


def run(N, B, aptonly, top=6):
    cfg = O.GwnetConfig(num_nodes=N, dropout=0.0, n_static_supports=0 if aptonly else 2, has_supports=not aptonly)
    gen = torch.Generator().manual_seed(0)
    sup = None if aptonly else O.synthetic_supports(N, 0.05, gen)
    x, _ = O.synthetic_batch(B, N, 12, 2, gen)
    x = torch.nn.functional.pad(x, (1, 0, 0, 0))
    torch.manual_seed(999)
    m = M.gwnet(dev, N, 0.0, supports=None if aptonly else [s.to(dev) for s in sup]).to(dev)
    st = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    m.train()
    out = m(x.to(dev))
    probe = torch.randn(out.shape, generator=gen)
    (out * probe.to(dev)).sum().backward()
    g1 = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    out2 = m(x.to(dev))
    (out2 * probe.to(dev)).sum().backward()
    rep = max(((p.grad - g1[k]).norm() / (g1[k].norm() + 1e-30)).item() for k, p in m.named_parameters() if p.grad is not None)
    s = {k: (v.clone().double() if v.is_floating_point() else v.clone()) for k, v in st.items()}
    pk = [k for k in s if not O.is_buffer(k)]
    for k in pk:
        s[k].requires_grad_(True)
    oo = O.forward(s, cfg, x.double(), [t.double() for t in sup] if sup else None, True)
    (oo * probe.double()).sum().backward()
    res = []
    for k in pk:
        if s[k].grad is None or "mlp.mlp.bias" in k:
            continue
        res.append((((g1[k].cpu().double() - s[k].grad).norm() / s[k].grad.norm()).item(), k))
    res.sort(reverse=True)
    print(f"N={N} B={B} aptonly={aptonly}: out rel {((out.detach().cpu().double() - oo.detach()).norm() / oo.detach().norm()).item():.2e} "
          f"run-to-run {rep:.2e}  worst: " + ", ".join(f"{k}={e:.1e}" for e, k in res[:top]), flush=True)


for N, B, apt in [(325, 16, True), (325, 16, False), (207, 16, True), (325, 64, True), (325, 8, True), (321, 16, True), (64, 16, True)]:
    run(N, B, apt)
