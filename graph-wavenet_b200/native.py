"""ctypes binding of the C ABI declared in ``include/gwnet_b200.h``.

The product path loads exactly one library: ``csrc/libgwnet_b200.so`` (nvcc, sm_100a).  There is
no CPU fallback -- :func:`get_lib` raises if the library is missing or if no CUDA device is
visible.  (``Lib`` can also wrap another build of the same ABI when given an explicit path; the
test-suite uses that for the host-emulation build of the kernels' index arithmetic.  Nothing in
this package does.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libgwnet_b200.so")

PREC_FP32, PREC_TF32, PREC_BF16, PREC_FP32X3 = 0, 1, 2, 3
DROPOUT_NONE, DROPOUT_MASK, DROPOUT_PHILOX = 0, 1, 2

c_float_p = C.POINTER(C.c_float)
c_void_pp = C.POINTER(C.c_void_p)


class GwnConfig(C.Structure):
    _fields_ = [
        ("batch", C.c_int), ("num_nodes", C.c_int), ("seq_len", C.c_int),
        ("in_dim", C.c_int), ("out_dim", C.c_int),
        ("residual_channels", C.c_int), ("dilation_channels", C.c_int),
        ("skip_channels", C.c_int), ("end_channels", C.c_int),
        ("kernel_size", C.c_int), ("blocks", C.c_int), ("layers", C.c_int),
        ("n_static_supports", C.c_int), ("gcn_bool", C.c_int), ("adaptive", C.c_int), ("gcn", C.c_int),
        ("order", C.c_int), ("apt_rank", C.c_int), ("precision", C.c_int),
        ("dropout", C.c_float), ("bn_eps", C.c_float), ("bn_momentum", C.c_float),
        ("dilation_base", C.c_int), ("per_sample_supports", C.c_int), ("adaptive_input", C.c_int),
    ]


class GwnForwardArgs(C.Structure):
    _fields_ = [
        ("params", c_void_pp), ("supports", c_void_pp), ("support_strides", C.POINTER(C.c_int64)),
        ("input", C.c_void_p), ("input_strides", C.c_int64 * 4), ("output", C.c_void_p),
        ("workspace", C.c_void_p), ("training", C.c_int), ("dropout_mode", C.c_int),
        ("keep_masks", c_void_pp), ("seed", C.c_uint64), ("stream", C.c_void_p), ("seed_device", C.c_void_p),
        ("apt_e1", C.c_void_p), ("apt_e2", C.c_void_p),
    ]


class GwnBackwardArgs(C.Structure):
    _fields_ = [
        ("params", c_void_pp), ("supports", c_void_pp), ("support_strides", C.POINTER(C.c_int64)),
        ("input", C.c_void_p), ("input_strides", C.c_int64 * 4), ("grad_output", C.c_void_p),
        ("workspace", C.c_void_p), ("scratch", C.c_void_p), ("grad_flat", C.c_void_p),
        ("grad_input", C.c_void_p), ("training", C.c_int), ("dropout_mode", C.c_int),
        ("keep_masks", c_void_pp), ("seed", C.c_uint64), ("stream", C.c_void_p), ("seed_device", C.c_void_p),
    ]


class GwnTrainArgs(C.Structure):
    _fields_ = [
        ("fwd", GwnForwardArgs), ("scratch", C.c_void_p), ("grad_flat", C.c_void_p), ("target", C.c_void_p),
        ("target_strides", C.c_int64 * 3), ("scaler_mean", C.c_float), ("scaler_std", C.c_float),
        ("ctrl", C.c_void_p), ("metrics", C.c_void_p),
    ]


class GwnAdamArgs(C.Structure):
    _fields_ = [
        ("param_flat", C.c_void_p), ("grad_flat", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
        ("live4", C.c_void_p), ("n", C.c_int64), ("hyper", C.c_void_p), ("ctrl", C.c_void_p), ("metrics", C.c_void_p),
        ("stream", C.c_void_p),
    ]


class GwnP2PArgs(C.Structure):
    _fields_ = [("base", C.c_void_p * 8), ("rank", C.c_int), ("world", C.c_int), ("sum_out", C.c_void_p)]


class GwnGcnDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("L", C.c_int), ("V", C.c_int), ("C", C.c_int), ("c_out", C.c_int),
        ("n_supports", C.c_int), ("order", C.c_int), ("precision", C.c_int), ("dropout_mode", C.c_int),
        ("dropout_p", C.c_float), ("seed", C.c_uint64), ("offset", C.c_uint64),
    ]


# every symbol include/gwnet_b200.h declares (tests check that the library exports them all)
EXPORTS = [
    "gwn_last_error", "gwn_abi_version", "gwn_launch_count", "gwn_device_info", "gwn_profile_begin", "gwn_profile_end", "gwn_permute4d",
    "gwn_node_contract", "gwn_node_contract_x3", "gwn_split_lo", "gwn_tc_error_flag", "gwn_tc_debug_buffer", "gwn_tc_debug_mode", "gwn_nconv_workspace_floats", "gwn_nconv_fwd", "gwn_nconv_bwd", "gwn_linear_fwd", "gwn_linear_bwd",
    "gwn_gcn_workspace_floats", "gwn_gcn_fwd", "gwn_gcn_bwd_scratch_floats", "gwn_gcn_bwd",
    "gwn_nconv2_fwd", "gwn_nconv2_bwd", "gwn_gcn2_fwd", "gwn_gcn2_bwd",
    "gwn_plan_create", "gwn_plan_destroy", "gwn_plan_workspace_bytes", "gwn_plan_param_count",
    "gwn_plan_param_info", "gwn_plan_out_len", "gwn_plan_debug_layout", "gwn_plan_forward", "gwn_plan_backward",
    "gwn_train_ctrl_bytes", "gwn_train_ctrl_init", "gwn_train_ctrl_read", "gwn_plan_train_fwd_bwd", "gwn_plan_eval_metrics", "gwn_adam_step",
    "gwn_p2p_header_bytes", "gwn_p2p_alloc", "gwn_p2p_open", "gwn_p2p_close", "gwn_p2p_free", "gwn_allreduce_adam_step",
]


class GwnError(RuntimeError):
    pass


def ptr_array(ptrs: Sequence[Optional[int]]):
    arr = (C.c_void_p * max(len(ptrs), 1))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


class Lib:
    """Typed view of one build of the ABI."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise GwnError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  gwnet_b200 has no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        d.gwn_last_error.restype = C.c_char_p
        d.gwn_abi_version.restype = C.c_int
        d.gwn_launch_count.argtypes = [C.c_int]
        d.gwn_launch_count.restype = C.c_longlong
        d.gwn_device_info.argtypes = [C.POINTER(C.c_int), C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int)]
        d.gwn_profile_end.argtypes = [C.c_char_p, C.c_int]
        i64p = C.POINTER(C.c_int64)
        d.gwn_permute4d.argtypes = [C.c_void_p, i64p, C.c_void_p, i64p, i64p, C.c_void_p]
        d.gwn_node_contract.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]
        d.gwn_node_contract_x3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]
        d.gwn_split_lo.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        d.gwn_tc_error_flag.argtypes = [C.c_int]
        d.gwn_tc_debug_buffer.argtypes = [C.c_void_p]
        d.gwn_tc_debug_buffer.restype = None
        d.gwn_tc_debug_mode.argtypes = [C.c_int]
        d.gwn_tc_debug_mode.restype = None
        d.gwn_nconv_workspace_floats.argtypes = [C.c_int] * 4
        d.gwn_nconv_workspace_floats.restype = C.c_size_t
        d.gwn_nconv_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]
        d.gwn_nconv_bwd.argtypes = ([C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
                                    + [C.c_int] * 5 + [C.c_void_p, C.c_void_p])
        d.gwn_linear_fwd.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_int, C.c_void_p]
        d.gwn_linear_bwd.argtypes = [C.c_void_p] * 6 + [C.c_int64, C.c_int, C.c_int, C.c_void_p]
        d.gwn_gcn_workspace_floats.argtypes = [C.POINTER(GwnGcnDesc), C.c_int]
        d.gwn_gcn_workspace_floats.restype = C.c_size_t
        d.gwn_gcn_fwd.argtypes = [C.POINTER(GwnGcnDesc), C.c_void_p, c_void_pp, i64p] + [C.c_void_p] * 7
        d.gwn_gcn_bwd_scratch_floats.argtypes = [C.POINTER(GwnGcnDesc)]
        d.gwn_gcn_bwd_scratch_floats.restype = C.c_size_t
        d.gwn_gcn_bwd.argtypes = ([C.POINTER(GwnGcnDesc), C.c_void_p, C.c_void_p, c_void_pp, i64p] + [C.c_void_p] * 6
                                  + [c_void_pp, i64p, C.c_void_p, C.c_void_p])
        d.gwn_nconv2_fwd.argtypes = ([C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p] + [C.c_int] * 5
                                     + [C.c_void_p, C.c_void_p])
        d.gwn_nconv2_bwd.argtypes = ([C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_int64] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p])
        d.gwn_gcn2_fwd.argtypes = [C.POINTER(GwnGcnDesc), C.c_void_p, c_void_pp, i64p, i64p] + [C.c_void_p] * 7
        d.gwn_gcn2_bwd.argtypes = ([C.POINTER(GwnGcnDesc), C.c_void_p, C.c_void_p, c_void_pp, i64p, i64p] + [C.c_void_p] * 6
                                   + [c_void_pp, i64p, i64p, C.c_void_p, C.c_void_p])
        d.gwn_plan_create.argtypes = [C.POINTER(GwnConfig), C.POINTER(C.c_void_p)]
        d.gwn_plan_destroy.argtypes = [C.c_void_p]
        d.gwn_plan_destroy.restype = None
        d.gwn_plan_workspace_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        d.gwn_plan_param_count.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
        d.gwn_plan_param_info.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64),
                                          C.POINTER(C.c_int64)]
        d.gwn_plan_out_len.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        d.gwn_plan_debug_layout.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        d.gwn_plan_forward.argtypes = [C.c_void_p, C.POINTER(GwnForwardArgs)]
        d.gwn_plan_backward.argtypes = [C.c_void_p, C.POINTER(GwnBackwardArgs)]
        d.gwn_train_ctrl_bytes.restype = C.c_size_t
        d.gwn_train_ctrl_init.argtypes = [C.c_void_p, C.c_uint64, C.c_int64]
        d.gwn_train_ctrl_read.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
        d.gwn_plan_train_fwd_bwd.argtypes = [C.c_void_p, C.POINTER(GwnTrainArgs)]
        d.gwn_plan_eval_metrics.argtypes = [C.c_void_p, C.POINTER(GwnTrainArgs)]
        d.gwn_adam_step.argtypes = [C.POINTER(GwnAdamArgs)]
        d.gwn_p2p_header_bytes.restype = C.c_size_t
        d.gwn_p2p_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
        d.gwn_p2p_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        d.gwn_p2p_close.argtypes = [C.c_void_p]
        d.gwn_p2p_free.argtypes = [C.c_void_p]
        d.gwn_allreduce_adam_step.argtypes = [C.POINTER(GwnAdamArgs), C.POINTER(GwnP2PArgs)]
        if d.gwn_abi_version() != 4:
            raise GwnError(f"{path}: ABI version {d.gwn_abi_version()} != 4")

    def check(self, status: int, what: str = ""):
        if status != 0:
            msg = self.dll.gwn_last_error()
            raise GwnError(f"{what or 'gwnet_b200'} failed (status {status}): {msg.decode() if msg else ''}")

    def device_info(self):
        n = C.c_int(0)
        sm = C.c_int(0)
        mj = C.c_int(0)
        mn = C.c_int(0)
        name = C.create_string_buffer(256)
        self.check(self.dll.gwn_device_info(C.byref(n), name, 256, C.byref(sm), C.byref(mj), C.byref(mn)), "device_info")
        return {"n_devices": n.value, "name": name.value.decode(), "sm_count": sm.value, "cc": (mj.value, mn.value)}


class Plan:
    """Owner of one ``gwn_plan`` handle."""

    def __init__(self, lib: Lib, cfg: GwnConfig):
        self.lib = lib
        self.cfg = cfg
        h = C.c_void_p()
        lib.check(lib.dll.gwn_plan_create(C.byref(cfg), C.byref(h)), "gwn_plan_create")
        self.handle = h
        f, b = C.c_size_t(0), C.c_size_t(0)
        lib.check(lib.dll.gwn_plan_workspace_bytes(h, C.byref(f), C.byref(b)))
        self.fwd_bytes, self.bwd_bytes = f.value, b.value
        n, g = C.c_int(0), C.c_int64(0)
        lib.check(lib.dll.gwn_plan_param_count(h, C.byref(n), C.byref(g)))
        self.n_entries, self.grad_floats = n.value, g.value
        self.names: List[str] = []
        self.grad_offsets: List[int] = []
        self.numels: List[int] = []
        buf = C.create_string_buffer(256)
        for i in range(self.n_entries):
            off, ne = C.c_int64(0), C.c_int64(0)
            lib.check(lib.dll.gwn_plan_param_info(h, i, buf, 256, C.byref(off), C.byref(ne)))
            self.names.append(buf.value.decode())
            self.grad_offsets.append(off.value)
            self.numels.append(ne.value)
        t, rf = C.c_int(0), C.c_int(0)
        lib.check(lib.dll.gwn_plan_out_len(h, C.byref(t), C.byref(rf)))
        self.t_out, self.receptive_field = t.value, rf.value

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.dll.gwn_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def forward(self, args: GwnForwardArgs):
        self.lib.check(self.lib.dll.gwn_plan_forward(self.handle, C.byref(args)), "gwn_plan_forward")

    def backward(self, args: GwnBackwardArgs):
        self.lib.check(self.lib.dll.gwn_plan_backward(self.handle, C.byref(args)), "gwn_plan_backward")


_LIB: Optional[Lib] = None


def get_lib() -> Lib:
    """The CUDA library, or an exception.  Never a substitute."""
    global _LIB
    if _LIB is None:
        lib = Lib(LIB_PATH)
        info = lib.device_info()
        if info["n_devices"] <= 0:
            raise GwnError("no CUDA device is visible: gwnet_b200 runs on sm_100a only and has no CPU fallback")
        _LIB = lib
    return _LIB
