"""Shared helpers for the parity tests (golden loading, comparators)."""
import json
import os

import numpy as np
import torch

from oracle import gwnet_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["dbl_adp", "aptonly", "static_only", "nogcn", "long_seq", "aptinit", "c32", "tr_long", "tr_c32"]
# cases recorded with the trainer's widths (skip = 8 nhid, end = 16 nhid): 3 engine.trainer.train steps + 1 eval of the real
# reference; tr_long / tr_c32 have inputs LONGER than the receptive field (T_out = 7)
TRAINER_CASES = ["dbl_adp", "aptonly", "tr_long", "tr_c32"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = O.GwnetConfig(**json.loads(str(z["cfg"])))
    rec = {"cfg": cfg}
    for k in z.files:
        if k == "cfg":
            continue
        v = z[k]
        rec[k] = torch.from_numpy(v) if v.ndim > 0 else torch.tensor(v.item())
    sup = [rec[f"support.{i}"] for i in range(cfg.n_static_supports)] if cfg.has_supports else None
    rec["supports"] = sup
    rec["state0"] = {k[len("state0/"):]: v.clone() for k, v in rec.items() if isinstance(k, str) and k.startswith("state0/")}
    # restore scalar dtype of num_batches_tracked
    for k in list(rec["state0"]):
        if k.endswith("num_batches_tracked"):
            rec["state0"][k] = torch.as_tensor(rec["state0"][k], dtype=torch.long).reshape(())
    return rec


def sub(rec, prefix):
    return {k[len(prefix):]: v for k, v in rec.items() if isinstance(k, str) and k.startswith(prefix)}


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    n = b.norm().item()
    return (a - b).norm().item() / (n if n > 0 else 1.0)


def assert_close_rel(a, b, tol, what="", floor=0.0):
    """Norm-relative comparison ||a-b|| <= tol*||b|| + floor (SURVEY.md App. C / G3:
    mathematically-zero gradients are compared through the absolute floor)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).norm().item()
    lim = tol * b.norm().item() + floor
    assert err <= lim, f"{what}: ||diff||={err:.3e} > {lim:.3e} (rel={err / max(b.norm().item(), 1e-30):.3e})"
