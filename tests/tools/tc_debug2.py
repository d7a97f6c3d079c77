import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build(); ge.load_package()
from graph_wavenet_b200 import native as NV
lib = NV.get_lib()
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=220, precision=2, sci_mode=False)
V = 32
x = torch.zeros(1, 4, V, 32)
for l in range(4):
    for k in range(V):
        x[0, l, k, :] = l * 1000 + k + torch.arange(32) / 100
S = torch.zeros(V, V)
for m in range(V):
    S[m, :] = 100 * m + torch.arange(V)
dbg = torch.full((16384 // 4 + 32 * 32 + 64,), -7.0, device=dev)
lib.dll.gwn_tc_debug_buffer(dbg.data_ptr())
y = torch.full(x.shape, float("nan"), device=dev)
st = torch.cuda.current_stream().cuda_stream
lib.check(lib.dll.gwn_node_contract(x.to(dev).data_ptr(), S.to(dev).data_ptr(), V, y.data_ptr(), 1, 4, V, 32, 1, st))
torch.cuda.synchronize()
d = dbg.cpu()
print("flag", lib.dll.gwn_tc_error_flag(1))
X = d[:4096].view(4, 32, 32)
print("X tile slab0 rows 0..9 (first 12 floats):\n", X[0, :10, :12])
print("X tile slab1 row 0:\n", X[1, 0, :12])
Bt = d[4096:4096 + 32 * 32].view(32, 32)
print("S tile rows 0..9 (first 12):\n", Bt[:10, :12])
print("y", y.cpu()[0, 0, :3, :6])
