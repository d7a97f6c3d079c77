"""GPU: check the tcgen05 node contraction (gwn_node_contract, TF32) against the fp32 FMA tier and torch fp64."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build(); ge.load_package()
from graph_wavenet_b200 import native as NV
lib = NV.get_lib()
dev = torch.device("cuda:0")


def run(B, L, V, C=32, iters=0):
    gen = torch.Generator().manual_seed(V * 7 + L)
    ld = (V + 3) // 4 * 4
    S = torch.zeros(V, ld)
    S[:, :V] = torch.softmax(torch.randn(V, V, generator=gen), dim=1)     # S[m][k]
    x = torch.randn(B, L, V, C, generator=gen)
    ref = torch.einsum("mk,blkc->blmc", S[:, :V].double(), x.double())
    Sd, xd = S.to(dev), x.to(dev)
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for name, prec in (("fp32", 0), ("tf32", 1)):
        y = torch.full((B, L, V, C), float("nan"), device=dev)
        lib.check(lib.dll.gwn_node_contract(xd.data_ptr(), Sd.data_ptr(), ld, y.data_ptr(), B, L, V, C, prec, st), name)
        torch.cuda.synchronize()
        flag = lib.dll.gwn_tc_error_flag(1)
        err = ((y.cpu().double() - ref).norm() / ref.norm()).item()
        nan = int(torch.isnan(y).sum().item())
        out[name] = (err, nan, flag)
    msg = f"B={B} L={L} V={V}: " + "  ".join(f"{k}: rel={v[0]:.2e} nan={v[1]} flag={v[2]}" for k, v in out.items())
    if iters:
        y = torch.empty((B, L, V, C), device=dev)
        for name, prec in (("fp32", 0), ("tf32", 1)):
            for _ in range(3):
                lib.dll.gwn_node_contract(xd.data_ptr(), Sd.data_ptr(), ld, y.data_ptr(), B, L, V, C, prec, st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                lib.dll.gwn_node_contract(xd.data_ptr(), Sd.data_ptr(), ld, y.data_ptr(), B, L, V, C, prec, st)
            e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1) / iters
            fl = 2.0 * B * L * C * V * V
            by = 2.0 * B * L * V * C * 4
            msg += f"  | {name}: {ms*1e3:.1f} us {fl/ms/1e9:.1f} TFLOP/s {by/ms/1e6:.0f} GB/s"
    print(msg, flush=True)
    return out


if __name__ == "__main__":
    run(1, 4, 32)
    run(1, 4, 64)
    run(2, 3, 207)
    run(64, 12, 207, iters=20)
    run(64, 1, 207, iters=20)
    run(16, 12, 325, iters=20)
    run(8, 12, 1024, iters=10)
    run(8, 24, 2048, iters=5)
