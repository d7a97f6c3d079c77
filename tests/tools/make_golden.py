"""Generate golden vectors from the REAL reference (sklin93/Graph-WaveNet).

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/tools/make_golden.py

The reference is imported unmodified through a stub-only shim:
missing third-party modules ``ipdb`` / ``matplotlib`` / ``nibabel`` are stubbed and
``nn.Conv1d`` is aliased to ``nn.Conv2d`` while ``gwnet`` is constructed (the fork
declares 2-D-kernel Conv1d layers, which modern torch refuses to run; the
parameters drawn are bit-identical).  For every case the script stores, as a
compressed ``tests/golden/<case>.npz``: the config, inputs, supports, the full
initial state_dict, the reference forward output, every parameter gradient of
``sum(out * probe)``, the BN buffers after that training-mode forward, and the
three metrics + updated parameters of reference ``engine.trainer.train`` steps.

It also checks the oracle restatement against the reference at the full METR-LA
and PEMS-BAY shapes and writes the observed differences (plus SHA-256 digests of
the seed-999 reference parameters) to ``tests/golden/fullsize_report.json``.
"""
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import gwnet_oracle as O  # noqa: E402


def load_reference():
    for n in ("ipdb", "nibabel"):
        sys.modules.setdefault(n, types.ModuleType(n))
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    sys.path.insert(0, REF)
    spec = importlib.util.spec_from_file_location("ref_model", os.path.join(REF, "model.py"))
    ref_model = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_model)
    saved = sys.modules.get("model")
    sys.modules["model"] = ref_model          # engine.py:2 does `from model import *`
    try:
        spec = importlib.util.spec_from_file_location("ref_engine", os.path.join(REF, "engine.py"))
        ref_engine = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_engine)
    finally:
        if saved is None:
            del sys.modules["model"]
        else:
            sys.modules["model"] = saved
    return ref_model, ref_engine


class conv1d_as_conv2d:
    def __enter__(self):
        self.c1 = nn.Conv1d
        nn.Conv1d = nn.Conv2d

    def __exit__(self, *a):
        nn.Conv1d = self.c1


class Scaler:
    def __init__(self, mean, std):
        self.mean, self.std = mean, std

    def inverse_transform(self, d):
        return d * self.std + self.mean


def build_ref_gwnet(ref_model, cfg: O.GwnetConfig, supports, aptinit=None):
    with conv1d_as_conv2d():
        return ref_model.gwnet("cpu", cfg.num_nodes, cfg.dropout,
                               supports=supports if cfg.has_supports else None,
                               gcn_bool=cfg.gcn_bool, addaptadj=cfg.addaptadj, aptinit=aptinit,
                               in_dim=cfg.in_dim, out_dim=cfg.out_dim,
                               residual_channels=cfg.residual_channels, dilation_channels=cfg.dilation_channels,
                               skip_channels=cfg.skip_channels, end_channels=cfg.end_channels,
                               kernel_size=cfg.kernel_size, blocks=cfg.blocks, layers=cfg.layers)


CASES = {
    # name: (cfg kwargs, batch, seq_len fed to trainer (before its +1 pad), seed)
    "dbl_adp": (dict(num_nodes=13, dropout=0.0, n_static_supports=2, residual_channels=8, dilation_channels=8,
                     skip_channels=64, end_channels=128, out_dim=12), 3, 12, 11),
    "aptonly": (dict(num_nodes=17, dropout=0.0, n_static_supports=0, has_supports=False, residual_channels=8,
                     dilation_channels=8, skip_channels=64, end_channels=128, out_dim=12), 2, 12, 12),
    "static_only": (dict(num_nodes=9, dropout=0.0, n_static_supports=2, addaptadj=False, residual_channels=8,
                         dilation_channels=8, skip_channels=16, end_channels=32, out_dim=12), 2, 12, 13),
    "nogcn": (dict(num_nodes=11, dropout=0.0, n_static_supports=0, has_supports=False, gcn_bool=False,
                   addaptadj=False, residual_channels=8, dilation_channels=8, skip_channels=16, end_channels=32,
                   out_dim=12), 2, 12, 14),
    "long_seq": (dict(num_nodes=10, dropout=0.0, n_static_supports=2, residual_channels=8, dilation_channels=8,
                      skip_channels=16, end_channels=32, out_dim=6, in_dim=3), 2, 20, 15),
    "aptinit": (dict(num_nodes=12, dropout=0.0, n_static_supports=2, residual_channels=8, dilation_channels=8,
                     skip_channels=16, end_channels=32, out_dim=12), 2, 12, 16),
    "c32": (dict(num_nodes=15, dropout=0.0, n_static_supports=2, residual_channels=32, dilation_channels=32,
                 skip_channels=64, end_channels=64, out_dim=12, blocks=2, layers=2), 2, 12, 17),
    # trainer-width cases (skip = 8 nhid, end = 16 nhid) whose input is LONGER than the receptive field (RF = 7): the
    # trainer's +1 pad (engine.py:44) is then a real extra column and T_out = 13 - 7 + 1 = 7 (fused.py)
    "tr_long": (dict(num_nodes=10, dropout=0.0, n_static_supports=2, residual_channels=8, dilation_channels=8,
                     skip_channels=64, end_channels=128, out_dim=12, blocks=2, layers=2), 3, 12, 18),
    "tr_c32": (dict(num_nodes=15, dropout=0.0, n_static_supports=2, residual_channels=32, dilation_channels=32,
                    skip_channels=256, end_channels=512, out_dim=12, blocks=2, layers=2), 2, 12, 19),
}


def make_case(name, ref_model, ref_engine):
    kw, B, T, seed = CASES[name]
    cfg = O.GwnetConfig(**kw)
    gen = torch.Generator().manual_seed(seed)
    supports = O.synthetic_supports(cfg.num_nodes, 0.3, gen)[: cfg.n_static_supports] if cfg.has_supports else None
    aptinit = supports[0] if name == "aptinit" else None
    x, y = O.synthetic_batch(B, cfg.num_nodes, T, cfg.in_dim, gen)
    torch.manual_seed(seed)
    model = build_ref_gwnet(ref_model, cfg, supports, aptinit)
    # oracle init from the same seed must equal the reference's
    torch.manual_seed(seed)
    ostate = O.init_state(cfg, aptinit)
    rstate = model.state_dict()
    assert list(ostate.keys()) == list(rstate.keys()), (name, "state_dict key order")
    for k in rstate:
        assert torch.equal(ostate[k], rstate[k]), (name, k)
    # randomise BN affine + running stats a little so they are exercised
    with torch.no_grad():
        for i in range(cfg.n_layers):
            model.bn[i].weight.copy_(1.0 + 0.2 * torch.randn(cfg.residual_channels, generator=gen))
            model.bn[i].bias.copy_(0.1 * torch.randn(cfg.residual_channels, generator=gen))
            model.bn[i].running_mean.copy_(0.1 * torch.randn(cfg.residual_channels, generator=gen))
            model.bn[i].running_var.copy_(1.0 + 0.3 * torch.rand(cfg.residual_channels, generator=gen))
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}

    rec = {"cfg": json.dumps(cfg.to_dict()), "batch": B, "seq": T}
    rec["x"] = x.contiguous().numpy()            # logical [B,F,N,T]
    rec["y"] = y.numpy()
    if supports is not None:
        for i, s in enumerate(supports):
            rec[f"support.{i}"] = s.numpy()
    if aptinit is not None:
        rec["aptinit"] = aptinit.numpy()
    for k, v in state0.items():
        rec["state0/" + k] = v.numpy()

    xin = torch.nn.functional.pad(x, (1, 0, 0, 0))
    # ---- eval-mode forward
    model.eval()
    with torch.no_grad():
        rec["out_eval"] = model(xin).numpy()
    # ---- train-mode forward + backward with a fixed probe
    model.train()
    model.zero_grad()
    xin_g = xin.clone().requires_grad_(True)
    out = model(xin_g)
    probe = torch.randn(out.shape, generator=gen)
    (out * probe).sum().backward()
    rec["out_train"] = out.detach().numpy()
    rec["probe"] = probe.numpy()
    rec["grad_input"] = xin_g.grad.numpy()
    for k, p in model.named_parameters():
        if p.grad is not None:
            rec["grad/" + k] = p.grad.numpy().copy()
    for k, v in model.state_dict().items():
        if O.is_buffer(k):
            rec["buf1/" + k] = v.numpy().copy()

    # ---- reference engine.trainer steps from state0
    torch.manual_seed(seed)
    with conv1d_as_conv2d():
        eng = ref_engine.trainer(Scaler(54.0, 20.0), cfg.in_dim, cfg.out_dim, cfg.num_nodes, cfg.residual_channels,
                                 cfg.dropout, 1e-3, 1e-4, "cpu", supports if cfg.has_supports else None,
                                 cfg.gcn_bool, cfg.addaptadj, aptinit, cfg.blocks, cfg.layers)
    if (cfg.skip_channels, cfg.end_channels) == (cfg.residual_channels * 8, cfg.residual_channels * 16):
        eng.model.load_state_dict(state0)
        metrics = []
        yy = y[:, :, : cfg.out_dim]
        for _ in range(3):
            metrics.append(eng.train(x, yy))
        metrics.append(eng.eval(x, yy))
        rec["trainer_metrics"] = np.asarray(metrics, dtype=np.float64)
        for k, v in eng.model.state_dict().items():
            rec["state3/" + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(f"[golden] {name}: {os.path.getsize(os.path.join(OUT, name + '.npz')) / 1024:.0f} KiB")


def fullsize_report(ref_model):
    report = {}
    shapes = {
        "metr-la": (O.GwnetConfig(num_nodes=207, dropout=0.0, n_static_supports=2), 64, 0.05),
        "pems-bay-aptonly": (O.GwnetConfig(num_nodes=325, dropout=0.0, n_static_supports=0, has_supports=False), 16, 0.05),
    }
    for name, (cfg, B, dens) in shapes.items():
        gen = torch.Generator().manual_seed(0)
        supports = O.synthetic_supports(cfg.num_nodes, dens, gen) if cfg.has_supports else None
        x, _ = O.synthetic_batch(B, cfg.num_nodes, 12, cfg.in_dim, gen)
        xin = torch.nn.functional.pad(x, (1, 0, 0, 0))
        torch.manual_seed(999)                     # train.py:47
        model = build_ref_gwnet(ref_model, cfg, supports)
        torch.manual_seed(999)
        ostate = O.init_state(cfg)
        digest = hashlib.sha256()
        for k, v in model.state_dict().items():
            assert torch.equal(v, ostate[k]), k
            digest.update(v.numpy().tobytes())
        model.train()
        out = model(xin)
        probe = torch.randn(out.shape, generator=gen)
        (out * probe).sum().backward()
        params = [k for k in ostate if not O.is_buffer(k)]
        for k in params:
            ostate[k].requires_grad_(True)
        oout = O.forward(ostate, cfg, xin, supports, True)
        (oout * probe).sum().backward()
        gd = 0.0
        gn = 0.0
        for k, p in model.named_parameters():
            if p.grad is None:
                assert ostate[k].grad is None, k
                continue
            gd += float((p.grad - ostate[k].grad).double().pow(2).sum())
            gn += float(p.grad.double().pow(2).sum())
        report[name] = {
            "batch": B,
            "state_sha256_seed999": digest.hexdigest(),
            "out_max_abs_diff": float((out - oout).abs().max()),
            "out_rel_l2": float((out - oout).norm() / out.norm()),
            "grad_global_rel_l2": (gd / gn) ** 0.5,
            "n_params_with_grad": sum(p.grad is not None for p in model.parameters()),
            "n_params": sum(1 for _ in model.parameters()),
            "out_first8": [float(v) for v in out.flatten()[:8]],
        }
        print("[fullsize]", name, report[name])
    with open(os.path.join(OUT, "fullsize_report.json"), "w") as f:
        json.dump(report, f, indent=1)


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    ref_model, ref_engine = load_reference()
    only = sys.argv[1:]          # optional: case names to (re)generate; default = everything + the full-size report
    for name in (only or CASES):
        make_case(name, ref_model, ref_engine)
    if not only:
        fullsize_report(ref_model)


if __name__ == "__main__":
    main()
