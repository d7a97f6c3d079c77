import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build(); ge.load_package()
from graph_wavenet_b200 import native as NV
lib = NV.get_lib()
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=220, precision=1, sci_mode=False)
st = torch.cuda.current_stream().cuda_stream
V = 128
gen = torch.Generator().manual_seed(0)
S = torch.randint(-3, 4, (V, V), generator=gen).float()
x = torch.randint(-3, 4, (1, 4, V, 32), generator=gen).float()
def call(mode):
    lib.dll.gwn_tc_debug_mode(mode)
    y = torch.full(x.shape, float("nan"), device=dev)
    lib.check(lib.dll.gwn_node_contract(x.to(dev).data_ptr(), S.to(dev).data_ptr(), V, y.data_ptr(), 1, 4, V, 32, 1, st))
    torch.cuda.synchronize()
    print(f"mode {mode} flag {lib.dll.gwn_tc_error_flag(1)}")
    return y.cpu()
y = call(1)    # y[0, slab, w, c] should be (32*slab + c)*1000 + w
exp = (32 * torch.arange(4)[:, None, None] + torch.arange(32)[None, None, :]) * 1000.0 + torch.arange(V)[None, :, None]
print("mode1 max abs err", (y[0] - exp).abs().max().item()); print(y[0, 1, :3, :5])
y = call(2)    # D[i, n] = sum_k S[i,k] S[n,k]; y[0, slab, n, c] = D[32*slab + c, n]
D = S @ S.t()
exp = D.view(4, 32, V).permute(0, 2, 1)
print("mode2 max abs err", (y[0] - exp).abs().max().item(), "|y|", y.norm().item(), "|exp|", exp.norm().item()); print(y[0, 0, :3, :6]); print(exp[0, :3, :6])
y = call(0)
ref = torch.einsum("mk,blkc->blmc", S, x)
print("mode0 max abs err", (y - ref).abs().max().item(), "|y|", y.norm().item())
lib.dll.gwn_tc_debug_mode(0)
