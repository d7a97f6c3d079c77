"""GPU probe: how does tcgen05.mma kind::tf32 reduce an fp32 operand to tf32 -- truncation of the low 13 mantissa
bits, or rounding?  (Decides how the 3xTF32 split computes its low part.)  Uses the node contraction with an
identity support: y = tf32(x) * 1.0 accumulated in fp32."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build(); ge.load_package()
from graph_wavenet_b200 import native as NV
lib = NV.get_lib()
dev = torch.device("cuda:0")
V, B, L, C = 64, 1, 4, 32
gen = torch.Generator().manual_seed(0)
x = (torch.randn(B, L, V, C, generator=gen) * 3).to(dev)
xi = x.view(torch.int32)
trunc = (xi & ~0x1FFF).view(torch.float32)
rna = ((xi + 0x1000) & ~0x1FFF).view(torch.float32)                       # round to nearest, ties away (cvt.rna.tf32.f32)
rne = ((xi + 0xFFF + ((xi >> 13) & 1)) & ~0x1FFF).view(torch.float32)     # round to nearest even
for which, S in (("S=I (x is the MN-major A operand)", torch.eye(V, device=dev)),):
    y = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    lib.check(lib.dll.gwn_node_contract(x.data_ptr(), S.contiguous().data_ptr(), V, y.data_ptr(), B, L, V, C, NV.PREC_TF32, st))
    torch.cuda.synchronize()
    print(which, "| equal to trunc:", bool(torch.equal(y, trunc)), "| rna:", bool(torch.equal(y, rna)), "| rne:", bool(torch.equal(y, rne)),
          "| max |y-x|/|x|:", float(((y - x).abs() / x.abs()).max()))
# the support is the K-major B operand: x = ones column selector, S carries the probe values
Sv = (torch.randn(V, V, generator=gen) * 3).to(dev)
X = torch.zeros(B, L, V, C, device=dev); X[:, :, 0, :] = 1.0            # y[m] = S[m,0]
y = torch.empty_like(X)
lib.check(lib.dll.gwn_node_contract(X.data_ptr(), Sv.data_ptr(), V, y.data_ptr(), B, L, V, C, NV.PREC_TF32, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
s0 = Sv[:, 0].contiguous(); si = s0.view(torch.int32)
got = y[0, 0, :, 0]
print("support as B operand | trunc:", bool(torch.equal(got, (si & ~0x1FFF).view(torch.float32))),
      "| rna:", bool(torch.equal(got, ((si + 0x1000) & ~0x1FFF).view(torch.float32))))
