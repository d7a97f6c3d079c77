"""One tiny gwnet forward+backward on cuda:0 through the C ABI, checked against the oracle."""
import sys

import torch

import __graft_entry__ as ge
from oracle import gwnet_oracle as O


def run_smoke():
    ge.load_package()
    from graph_wavenet_b200 import model as M
    dev = torch.device("cuda:0")
    cfg = O.GwnetConfig(num_nodes=23, dropout=0.0, n_static_supports=2, residual_channels=32, dilation_channels=32,
                        skip_channels=64, end_channels=128)
    gen = torch.Generator().manual_seed(3)
    sup = O.synthetic_supports(cfg.num_nodes, 0.3, gen)
    x, _ = O.synthetic_batch(4, cfg.num_nodes, 12, cfg.in_dim, gen)
    x = torch.nn.functional.pad(x, (1, 0, 0, 0))
    torch.manual_seed(1)
    m = M.gwnet(dev, cfg.num_nodes, 0.0, supports=[s.to(dev) for s in sup], residual_channels=32, dilation_channels=32,
                skip_channels=64, end_channels=128).to(dev)
    m.train()
    state = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    out = m(x.to(dev))
    probe = torch.randn(out.shape, generator=gen) * O.relu_safe_positions(state, cfg, x, sup, True)
    (out * probe.to(dev)).sum().backward()
    torch.cuda.synchronize()
    pk = [k for k in state if not O.is_buffer(k)]
    for k in pk:
        state[k].requires_grad_(True)
    oout = O.forward(state, cfg, x, sup, True)
    (oout * probe).sum().backward()
    err = (out.cpu() - oout.detach()).norm() / oout.detach().norm()
    gerr2 = gn2 = 0.0
    for k, p in m.named_parameters():
        if state[k].grad is None:
            assert p.grad is None, k
            continue
        gerr2 += float((p.grad.cpu() - state[k].grad).double().pow(2).sum())
        gn2 += float(state[k].grad.double().pow(2).sum())
    gerr = (gerr2 / gn2) ** 0.5
    print(f"[smoke] output rel-L2 {err:.3e}  gradient rel-L2 {gerr:.3e}", flush=True)
    assert err < 1e-4 and gerr < 1e-4, "smoke parity failed"


if __name__ == "__main__":
    run_smoke()
