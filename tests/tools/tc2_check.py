"""GPU: the CTA-pair node contraction (nconv_tc2.cuh, V > 256) -- correctness against fp64 on ragged shapes in both
tensor-core tiers, then throughput at N = 2048 / 4096.  GWNET_B200_NCONV_2CTA=0 times the one-CTA kernel instead.
usage: python tests/tools/tc2_check.py [perf_only]"""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build(); ge.load_package()
from graph_wavenet_b200 import native as NV
lib = NV.get_lib()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
mode = "1cta" if os.environ.get("GWNET_B200_NCONV_2CTA", "1") == "0" else "2cta"


def run(tier, x, S, Slo, ld, y, B, L, V):
    if tier == "tf32":
        lib.check(lib.dll.gwn_node_contract(x.data_ptr(), S.data_ptr(), ld, y.data_ptr(), B, L, V, 32, NV.PREC_TF32, st))
    else:
        lib.check(lib.dll.gwn_node_contract_x3(x.data_ptr(), S.data_ptr(), Slo.data_ptr(), ld, y.data_ptr(), B, L, V, 32, st))


def check(B, L, V):
    gen = torch.Generator().manual_seed(V + L)
    ld = (V + 3) // 4 * 4
    S = torch.zeros(V, ld)
    S[:, :V] = torch.softmax(torch.randn(V, V, generator=gen), dim=1) * (1 + torch.rand(V, V, generator=gen))
    x = torch.randn(B, L, V, 32, generator=gen)
    ref = torch.einsum("mk,blkc->blmc", S[:, :V].double(), x.double())
    Sd, xd = S.to(dev), x.to(dev)
    Slo = torch.empty_like(Sd)
    lib.check(lib.dll.gwn_split_lo(Sd.data_ptr(), Slo.data_ptr(), Sd.numel(), st))
    out = {}
    for tier in ("tf32", "fp32x3"):
        y = torch.full((B, L, V, 32), float("nan"), device=dev)
        run(tier, xd, Sd, Slo, ld, y, B, L, V)
        torch.cuda.synchronize()
        flag = lib.dll.gwn_tc_error_flag(1)
        err = float((y.double().cpu() - ref).norm() / ref.norm())
        out[tier] = (err, flag, bool(torch.isnan(y).any()))
    print(json.dumps({"check": mode, "B": B, "L": L, "V": V, **{k: {"rel_l2": v[0], "flag": v[1], "nan": v[2]} for k, v in out.items()}}), flush=True)
    return out["tf32"][0] < 2e-3 and out["fp32x3"][0] < 5e-5 and not any(v[1] or v[2] for v in out.values())


def perf(V, B, L, iters=10):
    gen = torch.Generator().manual_seed(V)
    ld = (V + 3) // 4 * 4
    S = torch.zeros(V, ld)
    S[:, :V] = torch.softmax(torch.randn(V, V, generator=gen), dim=1)
    S = S.to(dev).contiguous()
    Slo = torch.empty_like(S)
    lib.check(lib.dll.gwn_split_lo(S.data_ptr(), Slo.data_ptr(), S.numel(), st))
    x = torch.randn(B, L, V, 32, device=dev)
    y = torch.empty_like(x)
    for tier in ("tf32", "fp32x3"):
        for _ in range(3):
            run(tier, x, S, Slo, ld, y, B, L, V)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            run(tier, x, S, Slo, ld, y, B, L, V)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(json.dumps({"perf": mode, "V": V, "B": B, "L": L, "tier": tier, "us": round(ms * 1e3, 1),
                          "TFLOPs": round(2.0 * B * L * 32 * V * V / ms / 1e9, 1), "flag": lib.dll.gwn_tc_error_flag(1)}), flush=True)


if __name__ == "__main__":
    ok = True
    if not (len(sys.argv) > 1 and sys.argv[1] in ("perf_only", "ncu", "small", "ncu_small")):
        for B, L, V in [(2, 3, 300), (5, 1, 325), (1, 8, 512), (3, 7, 1000), (2, 12, 2048), (1, 1, 257)]:
            ok = check(B, L, V) and ok
        print("correctness:", "OK" if ok else "FAILED", flush=True)
    if len(sys.argv) > 1 and sys.argv[1] == "small":     # METR-LA-sized graphs: one-CTA vs pair kernel (GWNET_B200_NCONV_2CTA_MINV=0)
        perf(207, 64, 12, iters=20)
        perf(207, 64, 96, iters=10)
        perf(256, 64, 96, iters=10)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ncu_small":  # profiler run on the METR-LA-sized contraction
        perf(207, 64, 96, iters=2)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ncu":        # short run for the profiler: 5 launches per tier at config 4's layer-0 shape
        perf(2048, 64, 24, iters=2)
        sys.exit(0)
    if ok:
        perf(2048, 64, 24)
        perf(2048, 16, 24)
        perf(4096, 16, 24, iters=5)
    sys.exit(0 if ok else 1)
