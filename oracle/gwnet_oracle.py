"""CPU oracle for the Graph WaveNet forward/backward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``graph-wavenet_b200/`` may import this
file; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or as
the timed CPU baseline -- never as the product path.

This is a *restatement* in plain fp32 PyTorch (CPU) of the algorithm in the
reference's ``model.py`` / ``engine.py`` (sklin93/Graph-WaveNet), written as
stateless functions over an ordered ``state`` dict whose keys/shapes are the
reference ``gwnet.state_dict()``.  Each function cites the
reference lines it follows.

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference itself, generated in
the build container by ``tests/tools/make_golden.py`` (which imports
``/root/reference/model.py`` and ``engine.py`` through a stub-only shim) and
committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this
oracle against those vectors (forward, every gradient, BN buffers, a full
``trainer.train`` step).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- config
@dataclass
class GwnetConfig:
    """Constructor arguments of the reference ``gwnet`` (model.py:83-86)."""
    num_nodes: int
    dropout: float = 0.3
    n_static_supports: int = 2       # len(supports) when supports is not None
    has_supports: bool = True        # False <=> supports=None (aptonly)
    gcn_bool: bool = True
    addaptadj: bool = True
    in_dim: int = 2
    out_dim: int = 12
    residual_channels: int = 32
    dilation_channels: int = 32
    skip_channels: int = 256
    end_channels: int = 512
    kernel_size: int = 2
    blocks: int = 4
    layers: int = 2
    order: int = 2                   # gcn default (model.py:33); gwnet never overrides it

    @property
    def n_layers(self) -> int:
        return self.blocks * self.layers

    @property
    def adaptive(self) -> bool:
        return self.gcn_bool and self.addaptadj

    @property
    def supports_len(self) -> int:   # model.py:109-128
        n = self.n_static_supports if self.has_supports else 0
        return n + (1 if self.adaptive else 0)

    @property
    def gcn_active(self) -> bool:
        """model.py:225: ``gcn_bool and self.supports is not None``.  self.supports
        becomes [] (not None) whenever the adaptive branch ran (model.py:115-116)."""
        return self.gcn_bool and (self.has_supports or self.adaptive)

    def dilations(self) -> List[int]:  # model.py:130-153
        out = []
        for _ in range(self.blocks):
            d = 1
            for _ in range(self.layers):
                out.append(d)
                d *= 2
        return out

    @property
    def receptive_field(self) -> int:  # model.py:107,131,154-155,171
        rf = 1
        for _ in range(self.blocks):
            scope = self.kernel_size - 1
            for _ in range(self.layers):
                rf += scope
                scope *= 2
        return rf

    def to_dict(self):
        return asdict(self)


# --------------------------------------------------------------------------- init
def init_state(cfg: GwnetConfig, aptinit: Optional[torch.Tensor] = None) -> "OrderedDict[str, torch.Tensor]":
    """Draw parameters from the *current* torch CPU generator in the order the
    reference constructor consumes it (model.py:95-169): start_conv, nodevec1,
    nodevec2, then per layer filter/gate/residual/skip/(bn)/gcn.mlp, then
    end_conv_1, end_conv_2.  Returned in state_dict order (App. F)."""
    C, D = cfg.residual_channels, cfg.dilation_channels
    k = cfg.kernel_size
    drawn: Dict[str, torch.Tensor] = {}

    def conv(name, cin, cout, kw):
        m = nn.Conv2d(cin, cout, kernel_size=(1, kw))
        drawn[name + ".weight"] = m.weight.detach().clone()
        drawn[name + ".bias"] = m.bias.detach().clone()

    conv("start_conv", cfg.in_dim, C, 1)
    if cfg.adaptive:
        if aptinit is None:
            drawn["nodevec1"] = torch.randn(cfg.num_nodes, 10)
            drawn["nodevec2"] = torch.randn(10, cfg.num_nodes)
        else:  # model.py:123-127
            m, p, n = torch.svd(aptinit)
            drawn["nodevec1"] = torch.mm(m[:, :10], torch.diag(p[:10] ** 0.5))
            drawn["nodevec2"] = torch.mm(torch.diag(p[:10] ** 0.5), n[:, :10].t())
    gc_in = (cfg.order * cfg.supports_len + 1) * D
    for i in range(cfg.n_layers):
        conv(f"filter_convs.{i}", C, D, k)
        conv(f"gate_convs.{i}", C, D, k)
        conv(f"residual_convs.{i}", D, C, 1)
        conv(f"skip_convs.{i}", D, cfg.skip_channels, 1)
        drawn[f"bn.{i}.weight"] = torch.ones(C)
        drawn[f"bn.{i}.bias"] = torch.zeros(C)
        drawn[f"bn.{i}.running_mean"] = torch.zeros(C)
        drawn[f"bn.{i}.running_var"] = torch.ones(C)
        drawn[f"bn.{i}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
        if cfg.gcn_bool:
            conv(f"gconv.{i}.mlp.mlp", gc_in, C, 1)
    conv("end_conv_1", cfg.skip_channels, cfg.end_channels, 1)
    conv("end_conv_2", cfg.end_channels, cfg.out_dim, 1)

    order: List[str] = []
    if cfg.adaptive:
        order += ["nodevec1", "nodevec2"]
    for grp in ("filter_convs", "gate_convs", "residual_convs", "skip_convs"):
        for i in range(cfg.n_layers):
            order += [f"{grp}.{i}.weight", f"{grp}.{i}.bias"]
    for i in range(cfg.n_layers):
        order += [f"bn.{i}.{s}" for s in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")]
    if cfg.gcn_bool:
        for i in range(cfg.n_layers):
            order += [f"gconv.{i}.mlp.mlp.weight", f"gconv.{i}.mlp.mlp.bias"]
    for nme in ("start_conv", "end_conv_1", "end_conv_2"):
        order += [nme + ".weight", nme + ".bias"]
    return OrderedDict((kname, drawn[kname]) for kname in order)


BUFFER_SUFFIXES = ("running_mean", "running_var", "num_batches_tracked")


def is_buffer(key: str) -> bool:
    return key.endswith(BUFFER_SUFFIXES)


# --------------------------------------------------------------------------- operators
def nconv(x: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """model.py:12-14 -- y[n,c,w,l] = sum_v x[n,c,v,l] A[v,w], contiguous."""
    return torch.einsum("ncvl,vw->ncwl", x, A).contiguous()


def gcn(x, supports: Sequence[torch.Tensor], weight, bias, order: int = 2,
        dropout: float = 0.0, training: bool = False, keep_mask: Optional[torch.Tensor] = None):
    """model.py:41-55.  ``keep_mask`` (already scaled by 1/(1-p)) replaces the
    Bernoulli draw of F.dropout when given."""
    out = [x]
    for a in supports:
        x1 = nconv(x, a)
        out.append(x1)
        for _ in range(2, order + 1):
            x1 = nconv(x1, a)
            out.append(x1)
    h = torch.cat(out, dim=1)
    h = F.conv2d(h, weight, bias)
    if keep_mask is not None:
        return h * keep_mask
    return F.dropout(h, dropout, training=training)


def adaptive_adj(nodevec1, nodevec2):
    """model.py:187."""
    return F.softmax(F.relu(torch.mm(nodevec1, nodevec2)), dim=1)


# --------------------------------------------------------------------------- forward
def forward(state: Dict[str, torch.Tensor], cfg: GwnetConfig, inp: torch.Tensor,
            supports: Optional[Sequence[torch.Tensor]], training: bool,
            keep_masks: Optional[Sequence[Optional[torch.Tensor]]] = None,
            momentum: float = 0.1, eps: float = 1e-5, taps: Optional[dict] = None) -> torch.Tensor:
    """``gwnet.forward`` (model.py:175-241).  BN buffers in ``state`` are updated
    in place in training mode, as nn.BatchNorm2d does."""
    rf = cfg.receptive_field
    T = inp.size(3)
    x = F.pad(inp, (rf - T, 0, 0, 0)) if T < rf else inp             # :176-180
    x = F.conv2d(x, state["start_conv.weight"], state["start_conv.bias"])  # :181
    skip = None
    sup: Optional[List[torch.Tensor]] = None
    if cfg.gcn_active:
        sup = list(supports) if (supports is not None and cfg.has_supports) else []
        if cfg.adaptive:                                              # :185-188
            sup = sup + [adaptive_adj(state["nodevec1"], state["nodevec2"])]
    for i, d in enumerate(cfg.dilations()):                           # :192
        residual = x
        f = torch.tanh(F.conv2d(residual, state[f"filter_convs.{i}.weight"],
                                state[f"filter_convs.{i}.bias"], dilation=(1, d)))
        g = torch.sigmoid(F.conv2d(residual, state[f"gate_convs.{i}.weight"],
                                   state[f"gate_convs.{i}.bias"], dilation=(1, d)))
        x = f * g                                                     # :208-212
        s = F.conv2d(x, state[f"skip_convs.{i}.weight"], state[f"skip_convs.{i}.bias"])
        skip = s if skip is None else s + skip[:, :, :, -s.size(3):]  # :216-222
        if cfg.gcn_active:                                            # :225-230
            km = keep_masks[i] if keep_masks is not None else None
            x = gcn(x, sup, state[f"gconv.{i}.mlp.mlp.weight"], state[f"gconv.{i}.mlp.mlp.bias"],
                    cfg.order, cfg.dropout, training, km)
        else:                                                         # :232
            x = F.conv2d(x, state[f"residual_convs.{i}.weight"], state[f"residual_convs.{i}.bias"])
        x = x + residual[:, :, :, -x.size(3):]                        # :234
        rm, rv = state[f"bn.{i}.running_mean"], state[f"bn.{i}.running_var"]
        x = F.batch_norm(x, rm, rv, state[f"bn.{i}.weight"], state[f"bn.{i}.bias"],
                         training, momentum, eps)                     # :236
        if training:
            state[f"bn.{i}.num_batches_tracked"] += 1
    x = F.relu(skip)                                                  # :238
    e1 = F.conv2d(x, state["end_conv_1.weight"], state["end_conv_1.bias"])
    if taps is not None:            # ReLU inputs, for tests that must avoid the kink at 0
        taps["skip_pre"], taps["end1_pre"] = skip.detach(), e1.detach()
    x = F.relu(e1)
    return F.conv2d(x, state["end_conv_2.weight"], state["end_conv_2.bias"])


def relu_safe_positions(state, cfg: GwnetConfig, inp, supports, training: bool, tau: float = 1e-4,
                        keep_masks=None) -> torch.Tensor:
    """[B,1,N,T_out] 0/1 mask of output positions whose head ReLU inputs (model.py:238-239) all lie
    further than tau*rms from 0.  ReLU's derivative jumps at 0, so two correct fp32 implementations
    may disagree on the gate of an input that is ~1e-7 from it; gradient-parity tests multiply their
    probe by this mask so that such positions (whose influence is confined to themselves, the head
    being position-wise) carry no gradient in either implementation."""
    st = {k: v.detach().clone() for k, v in state.items()}
    taps = {}
    with torch.no_grad():
        forward(st, cfg, inp, supports, training, keep_masks, taps=taps)
    ok = None
    for z in taps.values():
        lim = tau * z.pow(2).mean().sqrt()
        m = (z.abs() > lim).all(dim=1, keepdim=True)
        ok = m if ok is None else (ok & m)
    return ok.float()


# --------------------------------------------------------------------------- losses
def _mask(labels, null_val):
    """Utils/util.py:511-517 (shared preamble of the masked losses)."""
    mask = (~torch.isnan(labels)) if (isinstance(null_val, float) and math.isnan(null_val)) else (labels != null_val)
    mask = mask.float()
    mask = mask / torch.mean(mask)
    return torch.where(torch.isnan(mask), torch.zeros_like(mask), mask)


def _masked_mean(loss, mask):
    loss = loss * mask
    loss = torch.where(torch.isnan(loss), torch.zeros_like(loss), loss)
    return torch.mean(loss)


def masked_mse(preds, labels, null_val=float("nan")):   # Utils/util.py:510-521
    return _masked_mean((preds - labels) ** 2, _mask(labels, null_val))


def masked_rmse(preds, labels, null_val=float("nan")):  # Utils/util.py:523-524
    return torch.sqrt(masked_mse(preds, labels, null_val))


def masked_mae(preds, labels, null_val=float("nan")):   # Utils/util.py:527-538
    return _masked_mean(torch.abs(preds - labels), _mask(labels, null_val))


def masked_mape(preds, labels, null_val=float("nan")):  # Utils/util.py:541-552
    return _masked_mean(torch.abs(preds - labels) / labels, _mask(labels, null_val))


# --------------------------------------------------------------------------- trainer step
class OracleTrainer:
    """Restatement of ``engine.trainer`` train/eval (engine.py:10-58,119-130) over the
    functional forward: Adam(lr, weight_decay as L2-in-grad), clip 5, masked MAE."""

    def __init__(self, cfg: GwnetConfig, state, supports, scaler_mean: float, scaler_std: float,
                 lrate: float = 1e-3, wdecay: float = 1e-4, clip: Optional[float] = 5.0):
        self.cfg = cfg
        self.state = state
        self.supports = supports
        self.mean, self.std = scaler_mean, scaler_std
        self.clip = clip
        self.params = [k for k in state if not is_buffer(k)]
        for k in self.params:
            state[k] = state[k].detach().clone().requires_grad_(True)
        self.optimizer = torch.optim.Adam([state[k] for k in self.params], lr=lrate, weight_decay=wdecay)

    def _predict(self, inp, training, keep_masks=None):
        inp = F.pad(inp, (1, 0, 0, 0))                                 # engine.py:44,121
        out = forward(self.state, self.cfg, inp, self.supports, training, keep_masks)
        out = out.transpose(1, 3)                                      # engine.py:46
        return out * self.std + self.mean                              # Utils/util.py:116-117

    def train(self, inp, real_val, keep_masks=None) -> Tuple[float, float, float]:
        self.optimizer.zero_grad()
        real = real_val.unsqueeze(1)
        predict = self._predict(inp, True, keep_masks)
        loss = masked_mae(predict, real, 0.0)
        loss.backward()
        if self.clip is not None:
            torch.nn.utils.clip_grad_norm_([self.state[k] for k in self.params], self.clip)
        self.optimizer.step()
        mape = masked_mape(predict, real, 0.0).item()
        rmse = masked_rmse(predict, real, 0.0).item()
        return loss.item(), mape, rmse

    def eval(self, inp, real_val) -> Tuple[float, float, float]:
        real = real_val.unsqueeze(1)
        predict = self._predict(inp, False)
        return (masked_mae(predict, real, 0.0).item(), masked_mape(predict, real, 0.0).item(),
                masked_rmse(predict, real, 0.0).item())


# --------------------------------------------------------------------------- synthetic workloads
def synthetic_supports(n: int, density: float, gen: torch.Generator) -> List[torch.Tensor]:
    """Row-stochastic forward/backward transition matrices with the value
    distribution of ``asym_adj`` on a sparse sensor graph (Utils/util.py:130-136,
    187-188)."""
    a = (torch.rand(n, n, generator=gen) < density).float() * torch.rand(n, n, generator=gen)
    a.fill_diagonal_(1.0)

    def rownorm(m):
        d = m.sum(1, keepdim=True)
        return torch.where(d > 0, m / d, torch.zeros_like(m))
    return [rownorm(a).contiguous(), rownorm(a.t().contiguous()).contiguous()]


def synthetic_batch(batch: int, n: int, seq: int, in_dim: int, gen: torch.Generator):
    """Input shaped like train.py:244-247 -- ``[B,T,N,F]`` host tensor viewed as
    ``[B,F,N,T]`` -- and target ``[B,N,T]`` with ~5 % exact zeros for the mask."""
    x = torch.randn(batch, seq, n, in_dim, generator=gen)
    if in_dim > 1:
        x[..., 1] = torch.rand(batch, seq, n, generator=gen)
    y = torch.rand(batch, n, seq, generator=gen) * 70.0
    y = torch.where(torch.rand(batch, n, seq, generator=gen) < 0.05, torch.zeros_like(y), y)
    return x.transpose(1, 3), y
