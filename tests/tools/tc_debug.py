import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build(); ge.load_package()
from graph_wavenet_b200 import native as NV
lib = NV.get_lib()
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=200, precision=2, sci_mode=False)

def go(V, S, x, tag):
    ld = (V + 3) // 4 * 4
    Sp = torch.zeros(V, ld); Sp[:, :V] = S
    B, L = x.shape[0], x.shape[1]
    y = torch.full(x.shape, float("nan"), device=dev)
    st = torch.cuda.current_stream().cuda_stream
    lib.check(lib.dll.gwn_node_contract(x.to(dev).data_ptr(), Sp.to(dev).data_ptr(), ld, y.data_ptr(), B, L, V, 32, 1, st))
    torch.cuda.synchronize()
    ref = torch.einsum("mk,blkc->blmc", S, x)
    y = y.cpu()
    print(f"--- {tag}: flag={lib.dll.gwn_tc_error_flag(1)} |y|={y.norm():.3f} |ref|={ref.norm():.3f} rel={(y-ref).norm()/ref.norm():.3e}")
    print("y[0,0,:8,:6]=\n", y[0, 0, :8, :6])
    print("ref[0,0,:8,:6]=\n", ref[0, 0, :8, :6])
    return y, ref

V = 32
k = torch.arange(V).float()
x = (k[None, None, :, None] + torch.arange(32).float()[None, None, None, :] / 100).expand(1, 4, V, 32).contiguous()
x = x + torch.arange(4).float()[None, :, None, None] * 100
go(V, torch.eye(V), x, "identity")
Ssh = torch.zeros(V, V); Ssh[torch.arange(V), (torch.arange(V) + 1) % V] = 1.0   # y[m] = x[m+1]
go(V, Ssh, x, "shift")
go(V, torch.ones(V, V), x, "ones")
y, ref = go(V, torch.eye(V), torch.ones(1, 4, V, 32), "identity-ones-x")
print("nonzero count", (y != 0).sum().item(), "unique", y.unique()[:10])
