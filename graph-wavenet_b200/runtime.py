"""Host-side driver of one ``gwn_plan``: argument marshalling for forward / backward.

PyTorch is used only for device memory (workspace, output and gradient buffers come from the
caching allocator) and for the current CUDA stream.  All compute happens behind the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import native as N


def make_config(*, batch, num_nodes, seq_len, in_dim, out_dim, residual_channels, dilation_channels, skip_channels,
                end_channels, kernel_size, blocks, layers, n_static_supports, gcn_bool, adaptive, gcn, order=2,
                apt_rank=10, precision=N.PREC_FP32, dropout=0.3, bn_eps=1e-5, bn_momentum=0.1, dilation_base=0,
                per_sample_supports=False, adaptive_input=False) -> N.GwnConfig:
    return N.GwnConfig(batch, num_nodes, seq_len, in_dim, out_dim, residual_channels, dilation_channels, skip_channels,
                       end_channels, kernel_size, blocks, layers, n_static_supports, int(gcn_bool), int(adaptive),
                       int(gcn), order, apt_rank, precision, float(dropout), float(bn_eps), float(bn_momentum),
                       int(dilation_base), int(per_sample_supports), int(adaptive_input))


def _stream_of(t: torch.Tensor) -> int:
    if t.is_cuda:
        return torch.cuda.current_stream(t.device).cuda_stream
    return 0


class ForwardCtx:
    """Everything backward needs from the matching forward call."""
    __slots__ = ("workspace", "input", "supports", "training", "dropout_mode", "masks", "seed", "keep", "runner")


class PlanRunner:
    def __init__(self, lib: N.Lib, cfg: N.GwnConfig):
        self.lib = lib
        self.plan = N.Plan(lib, cfg)
        self.cfg = cfg
        self.n_layers = cfg.blocks * cfg.layers
        self._scratch: Optional[torch.Tensor] = None

    # -- helpers
    def _param_table(self, tensors: Sequence[torch.Tensor]):
        if len(tensors) != self.plan.n_entries:
            raise N.GwnError(f"parameter table has {len(tensors)} entries, plan expects {self.plan.n_entries}")
        for t, name, ne in zip(tensors, self.plan.names, self.plan.numels):
            if t.numel() != ne or not t.is_contiguous():
                raise N.GwnError(f"parameter {name}: expected {ne} contiguous elements, got {tuple(t.shape)} "
                                 f"(contiguous={t.is_contiguous()})")
            want = torch.int64 if name.endswith("num_batches_tracked") else torch.float32
            if t.dtype != want:
                raise N.GwnError(f"parameter {name}: dtype {t.dtype}, expected {want}")
        return N.ptr_array([t.data_ptr() for t in tensors])

    def _supports(self, supports: Optional[Sequence[torch.Tensor]]):
        ns = self.cfg.n_static_supports
        sup = list(supports or [])[:ns]
        if len(sup) != ns:
            raise N.GwnError(f"expected {ns} static supports, got {len(sup)}")
        nn_ = self.cfg.num_nodes
        per = bool(self.cfg.per_sample_supports)
        want = (self.cfg.batch, nn_, nn_) if per else (nn_, nn_)
        for s in sup:
            if s.dtype != torch.float32 or tuple(s.shape) != want:
                raise N.GwnError(f"support must be fp32 {list(want)}, got {s.dtype} {tuple(s.shape)}")
        ptrs = N.ptr_array([s.data_ptr() for s in sup])
        k = 3 if per else 2
        strides = (C.c_int64 * max(k * ns, 1))()
        for i, s in enumerate(sup):
            for j in range(k):
                strides[k * i + j] = s.stride(j)
        return sup, ptrs, strides

    def forward(self, params: Sequence[torch.Tensor], supports, inp: torch.Tensor, training: bool,
                dropout_mode: int = N.DROPOUT_PHILOX, masks: Optional[Sequence[torch.Tensor]] = None, seed: int = 0,
                workspace: Optional[torch.Tensor] = None, apt=None):
        cfg = self.cfg
        if inp.dtype != torch.float32 or inp.dim() != 4 or tuple(inp.shape) != (cfg.batch, cfg.in_dim, cfg.num_nodes, cfg.seq_len):
            raise N.GwnError(f"input must be fp32 [{cfg.batch},{cfg.in_dim},{cfg.num_nodes},{cfg.seq_len}], got {inp.dtype} {tuple(inp.shape)}")
        dev = inp.device
        ptab = self._param_table(params)
        sup, sptrs, sstrides = self._supports(supports)
        out = torch.empty((cfg.batch, cfg.out_dim, cfg.num_nodes, self.plan.t_out), dtype=torch.float32, device=dev)
        if workspace is None:
            workspace = torch.empty(self.plan.fwd_bytes, dtype=torch.uint8, device=dev)
        a = N.GwnForwardArgs()
        a.params = ptab
        a.supports = sptrs
        a.support_strides = sstrides
        a.input = inp.data_ptr()
        for k in range(4):
            a.input_strides[k] = inp.stride(k)
        a.output = out.data_ptr()
        a.workspace = workspace.data_ptr()
        a.training = int(training)
        use_drop = training and cfg.dropout > 0 and cfg.gcn
        a.dropout_mode = dropout_mode if use_drop else N.DROPOUT_NONE
        mptrs = None
        if use_drop and dropout_mode == N.DROPOUT_MASK:
            if masks is None or len(masks) != self.n_layers:
                raise N.GwnError("DROPOUT_MASK needs one uint8 keep-mask per layer")
            for m in masks:
                if m.dtype != torch.uint8 or not m.is_contiguous():
                    raise N.GwnError("keep-masks must be contiguous uint8 in BLNC order")
            mptrs = N.ptr_array([m.data_ptr() for m in masks])
            a.keep_masks = mptrs
        a.seed = seed
        a.stream = _stream_of(inp)
        if cfg.adaptive_input:
            if apt is None or tuple(apt[0].shape) != (cfg.batch, cfg.num_nodes, cfg.apt_rank) or \
                    tuple(apt[1].shape) != (cfg.batch, cfg.apt_rank, cfg.num_nodes):
                raise N.GwnError("adaptive_input: node embeddings [B,N,rank] / [B,rank,N] required")
            a.apt_e1, a.apt_e2 = apt[0].data_ptr(), apt[1].data_ptr()
        self.plan.forward(a)
        ctx = ForwardCtx()
        ctx.workspace, ctx.input, ctx.supports = workspace, inp, sup
        ctx.training, ctx.dropout_mode, ctx.masks, ctx.seed = bool(training), a.dropout_mode, masks, seed
        ctx.keep = (ptab, sptrs, sstrides, mptrs, list(params), apt)
        return out, ctx

    def backward(self, ctx: ForwardCtx, params: Sequence[torch.Tensor], grad_out: torch.Tensor, need_input_grad: bool = False):
        cfg = self.cfg
        dev = ctx.input.device
        grad_out = grad_out.contiguous()
        if grad_out.dtype != torch.float32:
            raise N.GwnError("grad_output must be fp32")
        ptab = self._param_table(params)
        sup, sptrs, sstrides = self._supports(ctx.supports)
        grad_flat = torch.empty(self.plan.grad_floats, dtype=torch.float32, device=dev)
        if self._scratch is None or self._scratch.device != dev:
            self._scratch = torch.empty(self.plan.bwd_bytes, dtype=torch.uint8, device=dev)
        gin = torch.empty(ctx.input.shape, dtype=torch.float32, device=dev) if need_input_grad else None
        a = N.GwnBackwardArgs()
        a.params = ptab
        a.supports = sptrs
        a.support_strides = sstrides
        a.input = ctx.input.data_ptr()
        for k in range(4):
            a.input_strides[k] = ctx.input.stride(k)
        a.grad_output = grad_out.data_ptr()
        a.workspace = ctx.workspace.data_ptr()
        a.scratch = self._scratch.data_ptr()
        a.grad_flat = grad_flat.data_ptr()
        a.grad_input = gin.data_ptr() if gin is not None else None
        a.training = int(ctx.training)
        a.dropout_mode = ctx.dropout_mode
        mptrs = None
        if ctx.dropout_mode == N.DROPOUT_MASK:
            mptrs = N.ptr_array([m.data_ptr() for m in ctx.masks])
            a.keep_masks = mptrs
        a.seed = ctx.seed
        a.stream = _stream_of(grad_out)
        self.plan.backward(a)
        return grad_flat, gin

    def split_grads(self, grad_flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        out = {}
        for name, off, ne in zip(self.plan.names, self.plan.grad_offsets, self.plan.numels):
            if off >= 0:
                out[name] = grad_flat[off:off + ne]
        return out
