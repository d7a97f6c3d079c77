"""GPU: the fused hop-chain kernel (gcn_hops_fused.cuh, tf32 tier) against fp64 and against the two-launch chain.
Prints the relative error of both hop tensors (through the gcn operator's saved hops) and timings of the hop chain."""
import ctypes, json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build(); ge.load_package()
from graph_wavenet_b200 import native as NV
lib = NV.get_lib()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
mode = "two-launch" if os.environ.get("GWNET_B200_FUSED_HOPS", "1") == "0" else "fused"


def run(B, L, V, S, iters=0):
    gen = torch.Generator().manual_seed(V + L + S)
    x = torch.randn(B, L, V, 32, generator=gen)
    sup = [torch.softmax(torch.randn(V, V, generator=gen), dim=1) * (1 + torch.rand(V, V, generator=gen)) for _ in range(S)]
    W = torch.randn(32, (2 * S + 1) * 32, generator=gen) * 0.1
    b = torch.zeros(32)
    xd, Wd, bd = x.to(dev), W.to(dev), b.to(dev)
    supd = [s.to(dev).contiguous() for s in sup]
    hops = torch.full((2 * S, B, L, V, 32), float("nan"), device=dev)
    y = torch.empty(B, L, V, 32, device=dev)
    d = NV.GwnGcnDesc(B, L, V, 32, 32, S, 2, NV.PREC_TF32, NV.DROPOUT_NONE, 0.0, 0, 0)
    ws = torch.empty(int(lib.dll.gwn_gcn_workspace_floats(ctypes.byref(d), 0)), device=dev)
    sp = NV.ptr_array([s.data_ptr() for s in supd])
    lds = (ctypes.c_int64 * S)(*[V] * S)

    def call():
        lib.check(lib.dll.gwn_gcn_fwd(ctypes.byref(d), xd.data_ptr(), sp, lds, Wd.data_ptr(), bd.data_ptr(), None, hops.data_ptr(),
                                      y.data_ptr(), ws.data_ptr(), st), "gwn_gcn_fwd")
    call()
    torch.cuda.synchronize()
    flag = lib.dll.gwn_tc_error_flag(1)
    errs = []
    for s in range(S):
        h1 = torch.einsum("blvc,vw->blwc", x.double(), sup[s].double())
        h2 = torch.einsum("blvc,vw->blwc", h1, sup[s].double())
        e1 = float((hops[2 * s].double().cpu() - h1).norm() / h1.norm())
        e2 = float((hops[2 * s + 1].double().cpu() - h2).norm() / h2.norm())
        errs.append((round(e1, 6), round(e2, 6)))
    rec = {"mode": mode, "B": B, "L": L, "V": V, "S": S, "hop_rel_errors": errs, "nan": bool(torch.isnan(hops).any()), "flag": flag}
    if iters:
        for _ in range(3):
            call()
        e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            call()
        e1_.record(); e1_.synchronize()
        rec["us_per_gcn_fwd"] = round(e0.elapsed_time(e1_) / iters * 1e3, 1)
    print(json.dumps(rec), flush=True)
    return all(a < 3e-3 and b < 3e-3 for a, b in errs) and not rec["nan"] and flag == 0


ok = True
for B, L, V, S in [(2, 3, 53, 1), (3, 5, 207, 3), (1, 1, 16, 2), (5, 2, 256, 2), (2, 7, 100, 3), (9, 1, 207, 3)]:
    ok = run(B, L, V, S) and ok
print("fused hops:", "OK" if ok else "FAILED", flush=True)
if ok:
    run(64, 12, 207, 3, iters=20)
    run(64, 3, 207, 3, iters=20)
sys.exit(0 if ok else 1)
