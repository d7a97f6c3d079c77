"""CPU: the kernels' loader/epilogue functors and the plan orchestration, executed as serial host
loops (g++ -DGWN_HOST_EMU build of the very same sources), must reproduce the reference golden
vectors.  This is a test of index arithmetic and gradient formulas in the GPU-less container; the
parity tests proper are the ``-m gpu`` ones, which run the CUDA kernels."""
import os

import pytest
import torch

import __graft_entry__ as ge
from helpers import CASES, load_case, sub, assert_close_rel

pkg = ge.load_package()
from graph_wavenet_b200 import native as N          # noqa: E402
from graph_wavenet_b200.runtime import PlanRunner, make_config  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    path = ge.build_hostemu(os.path.join(os.path.dirname(__file__), "_hostemu"))
    return N.Lib(path)


def runner_for(emu, cfg, batch, seq_len, dropout=None):
    c = make_config(batch=batch, num_nodes=cfg.num_nodes, seq_len=seq_len, in_dim=cfg.in_dim, out_dim=cfg.out_dim,
                    residual_channels=cfg.residual_channels, dilation_channels=cfg.dilation_channels,
                    skip_channels=cfg.skip_channels, end_channels=cfg.end_channels, kernel_size=cfg.kernel_size,
                    blocks=cfg.blocks, layers=cfg.layers,
                    n_static_supports=cfg.n_static_supports if cfg.has_supports else 0, gcn_bool=cfg.gcn_bool,
                    adaptive=cfg.adaptive, gcn=cfg.gcn_active, order=cfg.order,
                    dropout=cfg.dropout if dropout is None else dropout)
    return PlanRunner(emu, c)


@pytest.mark.parametrize("name", CASES)
def test_emulated_plan_matches_reference(emu, name):
    rec = load_case(name)
    cfg = rec["cfg"]
    x = torch.nn.functional.pad(rec["x"], (1, 0, 0, 0))
    r = runner_for(emu, cfg, x.shape[0], x.shape[3])
    assert r.plan.names == list(rec["state0"].keys())
    params = [rec["state0"][k].clone().contiguous() for k in r.plan.names]
    # eval
    out, _ = r.forward(params, rec["supports"], x, training=False)
    assert_close_rel(out, rec["out_eval"], 2e-5, "eval output")
    # train fwd + bwd
    out, ctx = r.forward(params, rec["supports"], x, training=True)
    assert_close_rel(out, rec["out_train"], 2e-5, "train output")
    gflat, gin = r.backward(ctx, params, rec["probe"], need_input_grad=True)
    assert_close_rel(gin, rec["grad_input"], 1e-4, "grad input")
    grads = r.split_grads(gflat)
    ref = sub(rec, "grad/")
    gnorm = sum(float(g.double().pow(2).sum()) for g in ref.values()) ** 0.5
    for k, g in ref.items():
        assert_close_rel(grads[k].reshape(g.shape), g, 1e-4, "grad " + k, floor=2e-6 * gnorm)
    for k in grads:
        if k not in ref:   # dead parameters (G4): the plan must leave them at exactly zero
            assert float(grads[k].abs().max()) == 0.0, k
    bufs = sub(rec, "buf1/")
    for k, t in zip(r.plan.names, params):
        if k in bufs:
            assert_close_rel(t.float(), bufs[k].float(), 2e-5, "buffer " + k)


def test_emulated_dropout_mask_and_philox(emu):
    from oracle import gwnet_oracle as O
    rec = load_case("dbl_adp")
    cfg = rec["cfg"]
    cfg.dropout = 0.3
    x = torch.nn.functional.pad(rec["x"], (1, 0, 0, 0))
    r = runner_for(emu, cfg, x.shape[0], x.shape[3])
    params = [rec["state0"][k].clone().contiguous() for k in r.plan.names]
    gen = torch.Generator().manual_seed(5)
    B, N_, C = x.shape[0], cfg.num_nodes, cfg.residual_channels
    Ls, L = [], max(x.shape[3], cfg.receptive_field)
    for d in cfg.dilations():
        L -= d
        Ls.append(L)
    masks = [(torch.rand(B, l, N_, C, generator=gen) >= 0.3).to(torch.uint8) for l in Ls]     # BLNC
    out, ctx = r.forward(params, rec["supports"], x, training=True, dropout_mode=N.DROPOUT_MASK, masks=masks)
    state = {k: v.clone() for k, v in rec["state0"].items()}
    pk = [k for k in state if not O.is_buffer(k)]
    for k in pk:
        state[k].requires_grad_(True)
    keep = [m.permute(0, 3, 2, 1).float() / 0.7 for m in masks]                                # NCHW
    oout = O.forward(state, cfg, x, rec["supports"], True, keep)
    assert_close_rel(out, oout.detach(), 2e-5, "masked-dropout output")
    (oout * rec["probe"]).sum().backward()
    gflat, _ = r.backward(ctx, params, rec["probe"])
    grads = r.split_grads(gflat)
    gnorm = sum(float(state[k].grad.double().pow(2).sum()) for k in pk if state[k].grad is not None) ** 0.5
    for k in pk:
        if state[k].grad is not None:
            assert_close_rel(grads[k].reshape(state[k].shape), state[k].grad, 1e-4, "grad " + k, floor=2e-6 * gnorm)
    # Philox: deterministic per seed, different across seeds, keeps ~70 %
    o1, _ = r.forward([p.clone() for p in params], rec["supports"], x, training=True, dropout_mode=N.DROPOUT_PHILOX, seed=7)
    o2, _ = r.forward([p.clone() for p in params], rec["supports"], x, training=True, dropout_mode=N.DROPOUT_PHILOX, seed=7)
    o3, _ = r.forward([p.clone() for p in params], rec["supports"], x, training=True, dropout_mode=N.DROPOUT_PHILOX, seed=8)
    assert torch.equal(o1, o2) and not torch.equal(o1, o3)
