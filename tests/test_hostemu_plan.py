"""CPU: the kernels' loader/epilogue functors and the plan orchestration, executed as serial host
loops (g++ -DGWN_HOST_EMU build of the very same sources), must reproduce the reference golden
vectors.  This is a test of index arithmetic and gradient formulas in the GPU-less container; the
parity tests proper are the ``-m gpu`` ones, which run the CUDA kernels."""
import os

import pytest
import torch

import __graft_entry__ as ge
from helpers import CASES, TRAINER_CASES, load_case, sub, assert_close_rel

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


@pytest.mark.parametrize("name", TRAINER_CASES)
def test_emulated_fused_train_step_matches_reference_trainer(emu, name):
    """gwn_plan_train_fwd_bwd + gwn_adam_step (loss, metrics, clip, Adam as kernels) against 3 recorded
    ``engine.trainer.train`` steps of the real reference (metrics 1e-4, state 2e-3 -- see test_gpu_parity)."""
    import ctypes as C
    rec = load_case(name)
    cfg = rec["cfg"]
    # engine.py:44: the trainer left-pads by one zero column before gwnet.forward -- like fused.FusedStep, the plan is
    # built for T+1 and reads the zero column from the input buffer (tr_long / tr_c32: T+1 > RF, so it is a real column)
    x = torch.nn.functional.pad(rec["x"], (1, 0, 0, 0)).contiguous()
    y = rec["y"][:, :, : cfg.out_dim].contiguous()
    r = runner_for(emu, cfg, x.shape[0], x.shape[3])
    plan = r.plan
    assert plan.t_out == max(x.shape[3], cfg.receptive_field) - cfg.receptive_field + 1
    n = plan.grad_floats
    flat, grad = torch.zeros(n), torch.zeros(n)
    m, v = torch.zeros(n), torch.zeros(n)
    live4 = torch.zeros(n // 4, dtype=torch.uint8)
    last = cfg.blocks * cfg.layers - 1
    table = []
    for k, off, ne in zip(plan.names, plan.grad_offsets, plan.numels):
        t = rec["state0"][k].clone().contiguous()
        if off >= 0:
            flat[off:off + ne] = t.reshape(-1)
            t = flat[off:off + ne]
            dead = (k.startswith("residual_convs.") and cfg.gcn_active) or k.startswith(f"gconv.{last}.") or k.startswith(f"bn.{last}.")
            if not dead:
                live4[off // 4:(off + ne + 3) // 4] = 1
        table.append(t)
    ptab = N.ptr_array([t.data_ptr() for t in table])
    sup, sptrs, sstr = r._supports(rec["supports"])
    out = torch.zeros(x.shape[0], cfg.out_dim, cfg.num_nodes, plan.t_out)
    ws = torch.zeros(plan.fwd_bytes, dtype=torch.uint8)
    sc = torch.zeros(plan.bwd_bytes, dtype=torch.uint8)
    ctrl = torch.zeros(int(emu.dll.gwn_train_ctrl_bytes()), dtype=torch.uint8)
    emu.check(emu.dll.gwn_train_ctrl_init(ctrl.data_ptr(), 1234, 0))
    metrics = torch.zeros(4)
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 1e-4, 5.0, 1.0, 0.0])
    a = N.GwnTrainArgs()
    a.fwd.params, a.fwd.supports, a.fwd.support_strides = ptab, sptrs, sstr
    a.fwd.input = x.data_ptr()
    for k in range(4):
        a.fwd.input_strides[k] = x.stride(k)
    a.fwd.output, a.fwd.workspace, a.fwd.training = out.data_ptr(), ws.data_ptr(), 1
    a.fwd.dropout_mode = N.DROPOUT_NONE if cfg.dropout == 0 else N.DROPOUT_PHILOX
    a.scratch, a.grad_flat, a.target = sc.data_ptr(), grad.data_ptr(), y.data_ptr()
    for k in range(3):
        a.target_strides[k] = y.stride(k)
    a.scaler_mean, a.scaler_std, a.ctrl, a.metrics = 54.0, 20.0, ctrl.data_ptr(), metrics.data_ptr()
    ad = N.GwnAdamArgs()
    ad.param_flat, ad.grad_flat, ad.exp_avg, ad.exp_avg_sq = flat.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr()
    ad.live4, ad.n, ad.hyper, ad.ctrl, ad.metrics = live4.data_ptr(), n, hyper.data_ptr(), ctrl.data_ptr(), metrics.data_ptr()
    assert cfg.dropout == 0, "golden trainer cases are recorded without dropout"
    want = rec["trainer_metrics"].tolist()
    for step in range(3):
        emu.check(emu.dll.gwn_plan_train_fwd_bwd(plan.handle, C.byref(a)), "train_fwd_bwd")
        emu.check(emu.dll.gwn_adam_step(C.byref(ad)), "adam_step")
        for got, w in zip(metrics[:3].tolist(), want[step]):
            assert abs(got - w) <= 1e-4 * abs(w) + 1e-6, (step, metrics.tolist(), want[step])
    # trainer.eval after the three steps (engine.py:119-130): eval-mode forward + the three metrics
    emu.check(emu.dll.gwn_plan_eval_metrics(plan.handle, C.byref(a)), "eval_metrics")
    for got, w in zip(metrics[:3].tolist(), want[3]):
        assert abs(got - w) <= 1e-4 * abs(w) + 1e-6, ("eval", metrics.tolist(), want[3])
    seed, st = C.c_uint64(0), C.c_int64(0)
    emu.check(emu.dll.gwn_train_ctrl_read(ctrl.data_ptr(), C.byref(seed), C.byref(st)))
    assert st.value == 3 and seed.value != 1234
    for k, t in zip(plan.names, table):
        assert_close_rel(t.float().reshape(rec["state3/" + k].shape), rec["state3/" + k].float(), 2e-3, "state after 3 steps " + k,
                         floor=1e-5)


def _load_diffg():
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "diffg.npz"))
    rec = {k: torch.from_numpy(z[k]) if z[k].ndim > 0 else torch.tensor(z[k].item()) for k in z.files}
    state0 = {k[len("state0/"):]: v.clone() for k, v in rec.items() if k.startswith("state0/")}
    for k in state0:
        if k.endswith("num_batches_tracked"):
            state0[k] = torch.as_tensor(state0[k], dtype=torch.long).reshape(())
    return rec, state0


def test_emulated_per_sample_graph_plan_matches_reference_diff_G(emu):
    """gwnet_diff_G (model.py:244-407): per-sample supports, dilations 4/8, node embeddings drawn per forward -- the plan
    with per_sample_supports / adaptive_input / dilation_base against vectors of the real reference."""
    from oracle import diffg_oracle as DO
    rec, state0 = _load_diffg()
    x = rec["x"]
    B, F_, Nn, T = x.shape
    sup = [rec["support.0"], rec["support.1"]]
    c = make_config(batch=B, num_nodes=Nn, seq_len=T, in_dim=F_, out_dim=12, residual_channels=32, dilation_channels=32,
                    skip_channels=int(rec["cfg_skip"]), end_channels=int(rec["cfg_end"]), kernel_size=2, blocks=4, layers=2,
                    n_static_supports=2, gcn_bool=True, adaptive=False, gcn=True, order=2, dropout=0.0, dilation_base=4,
                    per_sample_supports=True, adaptive_input=True)
    r = PlanRunner(emu, c)
    assert r.plan.names == list(state0.keys())
    params = [state0[k].clone().contiguous() for k in r.plan.names]
    torch.manual_seed(int(rec["fwd_seed"]))
    apt = DO.draw_node_embeddings(B, Nn)
    out, ctx = r.forward(params, sup, x, training=True, apt=apt)
    assert_close_rel(out, rec["out_train"], 2e-5, "diff_G train output")
    gflat, gin = r.backward(ctx, params, rec["probe"], need_input_grad=True)
    assert_close_rel(gin, rec["grad_input"], 1e-4, "diff_G grad input")
    grads = r.split_grads(gflat)
    ref = sub(rec, "grad/")
    gnorm = sum(float(g.double().pow(2).sum()) for g in ref.values()) ** 0.5
    for k, g in ref.items():
        assert_close_rel(grads[k].reshape(g.shape), g, 1e-4, "diff_G grad " + k, floor=2e-6 * gnorm)
    for k, t in zip(r.plan.names, params):
        if "buf1/" + k in rec:
            assert_close_rel(t.float(), rec["buf1/" + k].float(), 2e-5, "diff_G buffer " + k)
    torch.manual_seed(int(rec["fwd_seed"]) + 1)
    apt = DO.draw_node_embeddings(B, Nn)
    out_e, _ = r.forward(params, sup, x, training=False, apt=apt)
    assert_close_rel(out_e, rec["out_eval"], 2e-5, "diff_G eval output")


def test_emulated_fused_step_null_labels_and_partial_mask(emu):
    """Masked-loss edge cases of Utils/util.py:510-552 through the loss kernels: an all-null target (mask empty: the
    reference turns the 0/0 mask into zeros, loss and metrics are 0, no gradient flows) and a partially null target,
    each against the oracle trainer."""
    import ctypes as C
    from oracle import gwnet_oracle as O
    rec = load_case("dbl_adp")
    cfg = rec["cfg"]
    x = rec["x"]
    r = runner_for(emu, cfg, x.shape[0], x.shape[3])
    plan = r.plan
    for case in ("all_null", "partial"):
        y = rec["y"][:, :, : cfg.out_dim].clone().contiguous()
        if case == "all_null":
            y.zero_()
        else:
            y[::2, :, ::3] = 0.0
        st_o = {k: v.clone() for k, v in rec["state0"].items()}
        tr = O.OracleTrainer(cfg, st_o, rec["supports"], 54.0, 20.0)
        want = tr.train(x, y)
        n = plan.grad_floats
        flat, grad, m, v = torch.zeros(n), torch.zeros(n), torch.zeros(n), torch.zeros(n)
        live4 = torch.zeros(n // 4, dtype=torch.uint8)
        last = cfg.blocks * cfg.layers - 1
        table = []
        for k, off, ne in zip(plan.names, plan.grad_offsets, plan.numels):
            t = rec["state0"][k].clone().contiguous()
            if off >= 0:
                flat[off:off + ne] = t.reshape(-1)
                t = flat[off:off + ne]
                dead = (k.startswith("residual_convs.") and cfg.gcn_active) or k.startswith(f"gconv.{last}.") or k.startswith(f"bn.{last}.")
                if not dead:
                    live4[off // 4:(off + ne + 3) // 4] = 1
            table.append(t)
        ptab = N.ptr_array([t.data_ptr() for t in table])
        sup, sptrs, sstr = r._supports(rec["supports"])
        out = torch.zeros(x.shape[0], cfg.out_dim, cfg.num_nodes, plan.t_out)
        ws = torch.zeros(plan.fwd_bytes, dtype=torch.uint8)
        sc = torch.zeros(plan.bwd_bytes, dtype=torch.uint8)
        ctrl = torch.zeros(int(emu.dll.gwn_train_ctrl_bytes()), dtype=torch.uint8)
        emu.check(emu.dll.gwn_train_ctrl_init(ctrl.data_ptr(), 1, 0))
        metrics = torch.full((4,), float("nan"))
        hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 1e-4, 5.0, 1.0, 0.0])
        a = N.GwnTrainArgs()
        a.fwd.params, a.fwd.supports, a.fwd.support_strides = ptab, sptrs, sstr
        a.fwd.input = x.data_ptr()
        for k in range(4):
            a.fwd.input_strides[k] = x.stride(k)
        a.fwd.output, a.fwd.workspace, a.fwd.training, a.fwd.dropout_mode = out.data_ptr(), ws.data_ptr(), 1, N.DROPOUT_NONE
        a.scratch, a.grad_flat, a.target = sc.data_ptr(), grad.data_ptr(), y.data_ptr()
        for k in range(3):
            a.target_strides[k] = y.stride(k)
        a.scaler_mean, a.scaler_std, a.ctrl, a.metrics = 54.0, 20.0, ctrl.data_ptr(), metrics.data_ptr()
        ad = N.GwnAdamArgs()
        ad.param_flat, ad.grad_flat, ad.exp_avg, ad.exp_avg_sq = flat.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr()
        ad.live4, ad.n, ad.hyper, ad.ctrl, ad.metrics = live4.data_ptr(), n, hyper.data_ptr(), ctrl.data_ptr(), metrics.data_ptr()
        emu.check(emu.dll.gwn_plan_train_fwd_bwd(plan.handle, C.byref(a)), "train_fwd_bwd")
        emu.check(emu.dll.gwn_adam_step(C.byref(ad)), "adam_step")
        assert not torch.isnan(metrics).any() and not torch.isnan(flat).any()
        for got, w in zip(metrics[:3].tolist(), want):
            assert abs(got - w) <= 1e-4 * abs(w) + 1e-6, (case, metrics.tolist(), want)
        if case == "all_null":
            assert metrics[:3].abs().max().item() == 0.0 and grad.abs().max().item() == 0.0
        for k, t in zip(plan.names, table):
            assert_close_rel(t.float().reshape(tr.state[k].shape), tr.state[k].detach().float(), 2e-3, f"{case}: state {k}", floor=1e-5)
