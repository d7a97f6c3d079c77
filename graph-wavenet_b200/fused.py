"""Fused optimisation step of ``engine.trainer.train`` (engine.py:41-58).

``FlatParams`` re-homes the model's parameters (and their ``.grad``) as views of two flat fp32 buffers laid out
like the plan's flat gradient buffer, ``FusedAdam`` is the ``torch.optim``-shaped front of ``gwn_adam_step``
(clip_grad_norm_ + Adam with L2 weight decay in one pass over those buffers) and ``FusedStep`` strings
``gwn_plan_train_fwd_bwd`` -> [NCCL all-reduce] -> ``gwn_adam_step`` -> metrics read-back into ONE CUDA graph.
PyTorch supplies memory, the stream, graph capture and the collective; all arithmetic is behind the C ABI.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import native as N


import contextlib


def _dev_ctx(dev):
    return torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else 0


def _sync(dev):
    if dev.type == "cuda":
        torch.cuda.current_stream(dev).synchronize()


def _pinned(t, dev):
    return t.pin_memory() if dev.type == "cuda" else t


class _DeviceMemory:
    """``torch.as_tensor`` view of device memory owned by the library (CUDA array interface, no copy)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class P2PComm:
    """Every rank's ``[flags | flat gradient]`` allocation, mapped into every process of the node (``gwn_p2p_alloc`` /
    ``gwn_p2p_open``; the 64-byte IPC handles travel through ``torch.distributed``).  ``grad`` is this rank's gradient
    buffer as a torch tensor: the backward pass writes it, the peers read it over NVLink inside
    ``gwn_allreduce_adam_step``."""

    def __init__(self, n_floats: int, dev, rank: int, world: int):
        import torch.distributed as dist
        lib = N.get_lib()
        self.lib, self.rank, self.world, self.n = lib, rank, world, int(n_floats)
        self.bases, self.own = [None] * world, None
        hdr = int(lib.dll.gwn_p2p_header_bytes())
        ok, err = 1, ""
        handle = C.create_string_buffer(64)
        with _dev_ctx(dev):
            base = C.c_void_p()
            if lib.dll.gwn_p2p_alloc(hdr + 4 * self.n, C.byref(base), handle) != 0:
                ok, err = 0, lib.dll.gwn_last_error().decode()
            else:
                self.own = base.value
            handles = [None] * world
            dist.all_gather_object(handles, (ok, bytes(handle.raw)))
            if all(h[0] for h in handles):
                for q in range(world):
                    if q == rank:
                        self.bases[q] = self.own
                        continue
                    p = C.c_void_p()
                    if lib.dll.gwn_p2p_open(handles[q][1], C.byref(p)) != 0:
                        ok, err = 0, lib.dll.gwn_last_error().decode()
                        break
                    self.bases[q] = p.value
            else:
                ok = 0
            flag = torch.tensor([ok], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # all ranks take the same path
            if int(flag.item()) == 0:
                self.close()
                raise N.GwnError("peer-memory gradient exchange unavailable" + (": " + err if err else ""))
            self.grad = torch.as_tensor(_DeviceMemory(self.own + hdr, self.n), device=dev)
            dist.barrier()

    def args(self, sum_out: torch.Tensor) -> N.GwnP2PArgs:
        a = N.GwnP2PArgs()
        for q in range(self.world):
            a.base[q] = self.bases[q]
        a.rank, a.world, a.sum_out = self.rank, self.world, sum_out.data_ptr()
        return a

    def close(self):
        for q, b in enumerate(self.bases):
            if b is not None and q != self.rank:
                self.lib.dll.gwn_p2p_close(b)
        self.bases = [None] * self.world
        if self.own is not None:
            self.lib.dll.gwn_p2p_free(self.own)
            self.own = None


def p2p_enabled() -> bool:
    return os.environ.get("GWNET_B200_P2P_ALLREDUCE", "1") != "0"


class FlatParams:
    """Parameters and gradients of one ``gwnet`` as views of flat buffers (offsets = the plan's gradient layout)."""

    def __init__(self, model, plan, p2p: Optional["P2PComm"] = None):
        named = list(model.named_parameters())
        offs = [o for o in plan.grad_offsets if o >= 0]
        nums = [n for n, o in zip(plan.numels, plan.grad_offsets) if o >= 0]
        if len(named) != len(offs):
            raise N.GwnError("gwnet_b200: parameter list does not match the native plan")
        dev = named[0][1].device
        self.n = plan.grad_floats
        if self.n % 4:
            raise N.GwnError("flat gradient buffer must be a multiple of 4 floats")
        self.param = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.p2p = p2p        # data parallel: the gradient buffer is the rank's peer-mapped allocation
        self.grad = p2p.grad if p2p is not None else torch.zeros(self.n, dtype=torch.float32, device=dev)
        if self.grad.numel() != self.n:
            raise N.GwnError("peer-memory gradient buffer does not match the plan's flat layout")
        live4 = torch.zeros(self.n // 4, dtype=torch.uint8)
        with torch.no_grad():
            for (name, p), off, ne in zip(named, offs, nums):
                if p.numel() != ne or p.dtype != torch.float32:
                    raise N.GwnError(f"parameter {name}: unexpected size/dtype for the native plan")
                self.param[off:off + ne].copy_(p.detach().reshape(-1))
                p.data = self.param[off:off + ne].view(p.shape)
                if p.requires_grad and model._live(name):
                    p.grad = self.grad[off:off + ne].view(p.shape)
                    live4[off // 4:(off + ne + 3) // 4] = 1
                else:
                    p.grad = None
        self.live4 = live4.to(dev)
        self.ptr0 = named[0][1].data_ptr()
        self.probe = named[0][1]
        model._entries = None

    def intact(self) -> bool:
        """False once something (``model.to``, ``p.data = ...``) moved the parameters out of the flat buffer."""
        return self.probe.data_ptr() == self.ptr0


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam(params, lr, weight_decay=wd)`` semantics (engine.py:33) executed by ``gwn_adam_step`` on the
    flat buffers; ``max_norm`` folds ``clip_grad_norm_`` (engine.py:53-54) into the same pass.  Hyper-parameters are
    read from ``param_groups[0]`` before every step, so schedulers that edit ``lr`` keep working."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat: Optional[FlatParams] = None
        self._hyper_cached = None
        self.max_norm = 0.0        # set by the trainer (self.clip)
        self.grad_scale = 1.0      # 1/world under data parallelism

    # -- binding to a model's flat buffers (done by the trainer on first use)
    def bind(self, flat: FlatParams, seed: int):
        lib = N.get_lib()
        dev = flat.param.device
        self._lib = lib
        self._flat = flat
        self.exp_avg = torch.zeros_like(flat.param)
        self.exp_avg_sq = torch.zeros_like(flat.param)
        self.ctrl = torch.zeros(int(lib.dll.gwn_train_ctrl_bytes()), dtype=torch.uint8, device=dev)
        self.metrics = torch.zeros(4, dtype=torch.float32, device=dev)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        self._hyper_host = _pinned(torch.zeros(8, dtype=torch.float32), dev)
        with _dev_ctx(dev):
            _sync(dev)
            lib.check(lib.dll.gwn_train_ctrl_init(self.ctrl.data_ptr(), int(seed) & (2 ** 64 - 1), 0), "gwn_train_ctrl_init")
        self._hyper_cached = None

        pend, self._pending = getattr(self, "_pending", None), None
        if pend is not None:      # a load_state_dict that arrived before the optimizer was bound
            self._restore(pend)

    @property
    def bound(self):
        return self._flat is not None

    # -- checkpointing: the moments and the step count live in flat device buffers / the control block, not in
    # Optimizer.state; state_dict()/load_state_dict() carry them under the extra key "fused" so that a resumed run
    # continues Adam where it stopped (torch.optim.Adam semantics)
    def state_dict(self):
        sd = super().state_dict()
        if self.bound:
            seed = C.c_uint64(0)
            step = C.c_int64(0)
            _sync(self._flat.param.device)
            self._lib.check(self._lib.dll.gwn_train_ctrl_read(self.ctrl.data_ptr(), C.byref(seed), C.byref(step)))
            sd["fused"] = {"exp_avg": self.exp_avg.detach().clone(), "exp_avg_sq": self.exp_avg_sq.detach().clone(),
                           "step": int(step.value), "seed": int(seed.value)}
        elif getattr(self, "_pending", None) is not None:
            sd["fused"] = self._pending
        return sd

    def load_state_dict(self, state_dict):
        sd = dict(state_dict)
        fused = sd.pop("fused", None)
        super().load_state_dict(sd)
        self._hyper_cached = None
        if fused is not None:
            if self.bound:
                self._restore(fused)
            else:
                self._pending = fused

    def _restore(self, fused):
        f = self._flat
        if fused["exp_avg"].numel() != f.n:
            raise N.GwnError("FusedAdam.load_state_dict: moment buffers do not match this model's flat layout")
        dev = f.param.device
        with torch.no_grad():
            self.exp_avg.copy_(fused["exp_avg"].to(dev))
            self.exp_avg_sq.copy_(fused["exp_avg_sq"].to(dev))
        with _dev_ctx(dev):
            _sync(dev)
            self._lib.check(self._lib.dll.gwn_train_ctrl_init(self.ctrl.data_ptr(), int(fused.get("seed", 0)) & (2 ** 64 - 1),
                                                              int(fused["step"])), "gwn_train_ctrl_init")

    def zero_grad(self, set_to_none: bool = False):
        """p.grad are views of the flat gradient buffer: zero the buffer and keep the views (set_to_none would detach
        them and leave p.grad None after fused steps)."""
        if self.bound:
            self._flat.grad.zero_()
        else:
            super().zero_grad(set_to_none=set_to_none)

    def step_count(self) -> int:
        seed, step = C.c_uint64(0), C.c_int64(0)
        self._lib.check(self._lib.dll.gwn_train_ctrl_read(self.ctrl.data_ptr(), C.byref(seed), C.byref(step)))
        return step.value

    def sync_hyper(self):
        g = self.param_groups[0]
        h = (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
             float(self.max_norm or 0.0), float(self.grad_scale), 0.0)
        if h != self._hyper_cached:
            self._hyper_host.copy_(torch.tensor(h, dtype=torch.float32))
            self.hyper.copy_(self._hyper_host, non_blocking=True)
            self._hyper_cached = h

    def adam_args(self, stream: int) -> N.GwnAdamArgs:
        f = self._flat
        a = N.GwnAdamArgs()
        a.param_flat, a.grad_flat = f.param.data_ptr(), f.grad.data_ptr()
        a.exp_avg, a.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        a.live4, a.n = f.live4.data_ptr(), f.n
        a.hyper, a.ctrl, a.metrics, a.stream = self.hyper.data_ptr(), self.ctrl.data_ptr(), self.metrics.data_ptr(), stream
        return a

    def launch(self):
        """Enqueue clip + Adam on the current stream (capturable)."""
        dev = self._flat.param.device
        a = self.adam_args(_stream(dev))
        self._lib.check(self._lib.dll.gwn_adam_step(C.byref(a)), "gwn_adam_step")

    def launch_p2p(self, sum_out: torch.Tensor):
        """Data parallel: gradient all-reduce over NVLink peer memory fused with the norm pass, then Adam (capturable)."""
        dev = self._flat.param.device
        a = self.adam_args(_stream(dev))
        p = self._flat.p2p.args(sum_out)
        self._lib.check(self._lib.dll.gwn_allreduce_adam_step(C.byref(a), C.byref(p)), "gwn_allreduce_adam_step")

    @torch.no_grad()
    def step(self, closure=None):
        """Eager use (``loss.backward(); optimizer.step()``): gradients must live in the bound flat buffer."""
        loss = closure() if closure is not None else None
        if not self.bound:
            raise N.GwnError("FusedAdam.step: optimizer is not bound to a gwnet (trainer.train does it on first use)")
        f = self._flat
        lo, hi = f.grad.data_ptr(), f.grad.data_ptr() + 4 * f.n
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None and not (lo <= p.grad.data_ptr() < hi):
                    raise N.GwnError("FusedAdam.step: a gradient lives outside the flat gradient buffer")
        self.sync_hyper()
        with _dev_ctx(f.param.device):
            self.launch()
        return loss


class FusedStep:
    """One captured ``trainer.train`` step for a fixed (batch, seq) shape."""

    def __init__(self, trainer, x: torch.Tensor, y: torch.Tensor, use_graph: bool):
        model, opt = trainer.model, trainer.optimizer
        lib = N.get_lib()
        self.lib, self.trainer = lib, trainer
        dev = x.device
        B, F, Nn, T = x.shape
        # engine.py:44 left-pads the input by one zero column BEFORE gwnet.forward, and model.py:176-180 pads to the
        # receptive field only when the padded length is still shorter: the plan is built for T+1 and reads the zero
        # column from the static input buffer (for T+1 <= RF the two pads coincide; for T >= RF they do not)
        runner = model._runner(B, T + 1)
        self.runner, self.plan = runner, runner.plan
        want_p2p = trainer.world > 1 and getattr(trainer, "p2p", False) and dev.type == "cuda"
        if model._flat is None or not model._flat.intact() or (want_p2p and model._flat.p2p is None):
            model._flat = FlatParams(model, runner.plan, trainer.p2p_comm(runner.plan.grad_floats, dev) if want_p2p else None)
            opt._flat = None
        if not opt.bound:
            opt.bind(model._flat, int(torch.randint(0, 2 ** 62, (1,)).item()))
        self.flat = model._flat
        # the peer-memory exchange: one kernel inside the step's single graph; else NCCL between two graphs (below)
        self.p2p = self.flat.p2p if trainer.world > 1 else None
        self.gsum = torch.empty(self.flat.n, dtype=torch.float32, device=dev) if self.p2p is not None else None
        cfg = runner.cfg
        # static buffers (the graph bakes their addresses)
        self.x = torch.zeros((B, T + 1, Nn, F), dtype=torch.float32, device=dev).permute(0, 3, 2, 1)   # loader layout [B,T,N,F]
        self.x_in = self.x[:, :, :, 1:]      # column 0 stays zero: the trainer's pad
        self.y = torch.empty(tuple(y.shape), dtype=torch.float32, device=dev)
        self.out = torch.empty((B, cfg.out_dim, Nn, self.plan.t_out), dtype=torch.float32, device=dev)
        self.workspace = torch.empty(self.plan.fwd_bytes, dtype=torch.uint8, device=dev)
        self.scratch = torch.empty(self.plan.bwd_bytes, dtype=torch.uint8, device=dev)
        self.metrics_host = _pinned(torch.zeros(4, dtype=torch.float32), dev)
        self.table = [t.detach() for t in model._table()]
        self.ptab = runner._param_table(self.table)
        self.sup, self.sptrs, self.sstrides = runner._supports(model.supports)
        if y.dim() != 3 or y.shape[0] != B or y.shape[1] != Nn or y.shape[2] != cfg.out_dim:
            raise N.GwnError(f"real_val must be [B={B}, N={Nn}, out_dim={cfg.out_dim}], got {tuple(y.shape)}")
        self.use_dropout = cfg.dropout > 0 and cfg.gcn
        self.masks = model._dropout_masks
        self.mptrs = None
        self.graph = None
        self.use_graph = use_graph = bool(use_graph and dev.type == "cuda")
        self.key_ptrs = self._ptr_key()
        self.graph_tail = None
        # Under data parallelism the step is captured as TWO graphs with the gradient all-reduce launched between them
        # by torch.distributed (NCCL's own stream-ordering, no host sync): capturing the collective itself depends on the
        # NCCL watchdog tolerating stream capture.  GWNET_B200_NCCL_IN_GRAPH=1 captures it into a single graph instead.
        self.split = trainer.world > 1 and self.p2p is None and os.environ.get("GWNET_B200_NCCL_IN_GRAPH", "0") != "1"
        if use_graph:
            if trainer.world > 1 and not self.split and self.p2p is None:     # NCCL must have built its communicator before capture
                import torch.distributed as dist
                dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize(dev)
            if self.split:
                g, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue_fwd_bwd()
                with torch.cuda.graph(g2):
                    self._enqueue_tail()
                self.graph, self.graph_tail = g, g2
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue()
                self.graph = g

    def _ptr_key(self):
        return (self.table[0].data_ptr(), self.table[-1].data_ptr(), tuple(s.data_ptr() for s in self.sup),
                self.trainer.world, id(self.trainer.model._dropout_masks))

    def valid(self) -> bool:
        m = self.trainer.model
        return (self.flat is m._flat and self.flat.intact() and m._entries is not None
                and self.key_ptrs[2] == tuple(s.data_ptr() for s in (m.supports or [])[:len(self.sup)])
                and self.key_ptrs[3] == self.trainer.world and self.key_ptrs[4] == id(m._dropout_masks))

    def _train_args(self, stream: int) -> N.GwnTrainArgs:
        a = N.GwnTrainArgs()
        f = a.fwd
        f.params, f.supports, f.support_strides = self.ptab, self.sptrs, self.sstrides
        f.input = self.x.data_ptr()
        for k in range(4):
            f.input_strides[k] = self.x.stride(k)
        f.output, f.workspace, f.training = self.out.data_ptr(), self.workspace.data_ptr(), 1
        f.dropout_mode = N.DROPOUT_NONE
        if self.use_dropout:
            if self.masks is not None:
                f.dropout_mode = N.DROPOUT_MASK
                self.mptrs = N.ptr_array([m.data_ptr() for m in self.masks])
                f.keep_masks = self.mptrs
            else:
                f.dropout_mode = N.DROPOUT_PHILOX
        f.stream = stream
        opt = self.trainer.optimizer
        a.scratch, a.grad_flat, a.target = self.scratch.data_ptr(), self.flat.grad.data_ptr(), self.y.data_ptr()
        for k in range(3):
            a.target_strides[k] = self.y.stride(k)
        a.scaler_mean, a.scaler_std = float(self.trainer.scaler.mean), float(self.trainer.scaler.std)
        a.ctrl, a.metrics = opt.ctrl.data_ptr(), opt.metrics.data_ptr()
        return a

    def _enqueue_fwd_bwd(self):
        dev = self.x.device
        with _dev_ctx(dev):
            a = self._train_args(_stream(dev))
            self.lib.check(self.lib.dll.gwn_plan_train_fwd_bwd(self.plan.handle, C.byref(a)), "gwn_plan_train_fwd_bwd")

    def _allreduce(self):
        import torch.distributed as dist
        dist.all_reduce(self.flat.grad)          # ONE collective per step; 1/world is folded into Adam

    def _enqueue_tail(self):
        tr = self.trainer
        with _dev_ctx(self.x.device):
            tr.optimizer.launch()
            self.metrics_host.copy_(tr.optimizer.metrics, non_blocking=True)

    def _enqueue(self):
        self._enqueue_fwd_bwd()
        if self.p2p is not None:
            tr = self.trainer
            with _dev_ctx(self.x.device):
                tr.optimizer.launch_p2p(self.gsum)
                self.metrics_host.copy_(tr.optimizer.metrics, non_blocking=True)
            return
        if self.trainer.world > 1:
            self._allreduce()
        self._enqueue_tail()

    def run(self, x: torch.Tensor, y: torch.Tensor):
        tr = self.trainer
        opt = tr.optimizer
        opt.max_norm = float(tr.clip) if tr.clip is not None else 0.0
        opt.grad_scale = 1.0 / tr.world
        opt.sync_hyper()
        self.x_in.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        if self.graph is not None and self.graph_tail is not None:
            self.graph.replay()
            self._allreduce()
            self.graph_tail.replay()
        elif self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        _sync(self.x.device)      # the step's single host sync
        tr.model._last_grad_flat = self.flat.grad
        m = self.metrics_host
        return float(m[0]), float(m[1]), float(m[2])


class FusedEval:
    """``trainer.eval`` (engine.py:119-130) for a fixed shape: forward in eval mode + the three masked metrics as one
    captured launch sequence, one host sync."""

    def __init__(self, trainer, x: torch.Tensor, y: torch.Tensor, use_graph: bool, workspace: Optional[torch.Tensor] = None):
        model = trainer.model
        lib = N.get_lib()
        self.lib, self.trainer = lib, trainer
        dev = x.device
        B, F, Nn, T = x.shape
        runner = model._runner(B, T + 1)     # engine.py:121: the +1 left pad is part of the network input (see FusedStep)
        self.runner, self.plan = runner, runner.plan
        cfg = runner.cfg
        if y.dim() != 3 or y.shape[0] != B or y.shape[1] != Nn or y.shape[2] != cfg.out_dim:
            raise N.GwnError(f"real_val must be [B={B}, N={Nn}, out_dim={cfg.out_dim}], got {tuple(y.shape)}")
        self.x = torch.zeros((B, T + 1, Nn, F), dtype=torch.float32, device=dev).permute(0, 3, 2, 1)
        self.x_in = self.x[:, :, :, 1:]
        self.y = torch.empty(tuple(y.shape), dtype=torch.float32, device=dev)
        self.out = torch.empty((B, cfg.out_dim, Nn, self.plan.t_out), dtype=torch.float32, device=dev)
        self.workspace = workspace if workspace is not None else torch.empty(self.plan.fwd_bytes, dtype=torch.uint8, device=dev)
        self.ctrl = torch.zeros(int(lib.dll.gwn_train_ctrl_bytes()), dtype=torch.uint8, device=dev)
        self.metrics = torch.zeros(4, dtype=torch.float32, device=dev)
        self.metrics_host = _pinned(torch.zeros(4, dtype=torch.float32), dev)
        self.table = [t.detach() for t in model._table()]
        self.ptab = runner._param_table(self.table)
        self.sup, self.sptrs, self.sstrides = runner._supports(model.supports)
        self.key = (self.table[0].data_ptr(), self.table[-1].data_ptr(), tuple(s.data_ptr() for s in self.sup))
        self.graph = None
        if use_graph and dev.type == "cuda":
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self.graph = g

    def valid(self) -> bool:
        m = self.trainer.model
        if m._entries is None:
            return False
        t = m._entries
        return self.key == (t[0].data_ptr(), t[-1].data_ptr(), tuple(s.data_ptr() for s in (m.supports or [])[:len(self.sup)]))

    def _enqueue(self):
        dev = self.x.device
        with _dev_ctx(dev):
            a = N.GwnTrainArgs()
            f = a.fwd
            f.params, f.supports, f.support_strides = self.ptab, self.sptrs, self.sstrides
            f.input = self.x.data_ptr()
            for k in range(4):
                f.input_strides[k] = self.x.stride(k)
            f.output, f.workspace, f.training, f.dropout_mode = self.out.data_ptr(), self.workspace.data_ptr(), 0, N.DROPOUT_NONE
            f.stream = _stream(dev)
            a.target = self.y.data_ptr()
            for k in range(3):
                a.target_strides[k] = self.y.stride(k)
            a.scaler_mean, a.scaler_std = float(self.trainer.scaler.mean), float(self.trainer.scaler.std)
            a.ctrl, a.metrics = self.ctrl.data_ptr(), self.metrics.data_ptr()
            self.lib.check(self.lib.dll.gwn_plan_eval_metrics(self.plan.handle, C.byref(a)), "gwn_plan_eval_metrics")
            self.metrics_host.copy_(self.metrics, non_blocking=True)

    def run(self, x: torch.Tensor, y: torch.Tensor):
        self.x_in.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        _sync(self.x.device)
        m = self.metrics_host
        return float(m[0]), float(m[1]), float(m[2])


def fused_enabled() -> bool:
    return os.environ.get("GWNET_B200_FUSED_STEP", "1") != "0"


def graph_enabled() -> bool:
    return os.environ.get("GWNET_B200_GRAPH", "1") != "0"
