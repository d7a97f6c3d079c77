"""CPU: the CUDA library builds, loads, exports every symbol of include/gwnet_b200.h, and refuses
to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import __graft_entry__ as ge

ge.build()
ge.load_package()
from graph_wavenet_b200 import native as N            # noqa: E402
from graph_wavenet_b200.runtime import make_config    # noqa: E402


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ge.ROOT, "include", "gwnet_b200.h")).read()
    declared = set(re.findall(r"^(?:long long|int|void|size_t|const char\*)\s+(gwn_[a-z0-9_]+)\s*\(", hdr, re.M))
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    dll = ctypes.CDLL(N.LIB_PATH)
    for sym in declared:
        assert hasattr(dll, sym), sym


def test_plan_metadata_without_device():
    lib = N.Lib(N.LIB_PATH)
    cfg = make_config(batch=64, num_nodes=207, seq_len=13, in_dim=2, out_dim=12, residual_channels=32,
                      dilation_channels=32, skip_channels=256, end_channels=512, kernel_size=2, blocks=4, layers=2,
                      n_static_supports=2, gcn_bool=1, adaptive=1, gcn=1)
    plan = N.Plan(lib, cfg)
    assert plan.n_entries == 128 and plan.t_out == 1 and plan.receptive_field == 13
    assert sum(n for n, o in zip(plan.numels, plan.grad_offsets) if o >= 0) == 309400
    assert plan.names[0] == "nodevec1" and plan.names[-1] == "end_conv_2.bias"


def test_bad_config_is_rejected_with_message():
    lib = N.Lib(N.LIB_PATH)
    cfg = make_config(batch=1, num_nodes=5, seq_len=13, in_dim=2, out_dim=12, residual_channels=30,
                      dilation_channels=32, skip_channels=256, end_channels=512, kernel_size=2, blocks=4, layers=2,
                      n_static_supports=2, gcn_bool=1, adaptive=1, gcn=1)
    with pytest.raises(N.GwnError, match="multiples of 4"):
        N.Plan(lib, cfg)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(N.GwnError, match="no CUDA device"):
        N.get_lib()
    from graph_wavenet_b200 import model as M
    m = M.gwnet("cpu", 7, 0.0, supports=[torch.eye(7)])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 2, 7, 13))
    lib = N.Lib(N.LIB_PATH)
    x = torch.zeros(1, 1, 4, 4)
    st = lib.dll.gwn_nconv_fwd(x.data_ptr(), x.data_ptr(), 4, x.data_ptr(), 1, 1, 4, 4, 0, None, None)
    assert st == 10004
