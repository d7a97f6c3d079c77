"""GPU: the other BASELINE.json configurations (not bench lines -- context numbers for DESIGN.md): full trainer.train step
(one CUDA graph) per tier at PEMS-BAY aptonly, CRASH shapes and the large-graph stress shape, plus the node contraction
alone at N = 2048 / 4096.  usage: python tests/tools/configs_bench.py [quick]"""
import json, os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
from oracle import gwnet_oracle as O
ge.build(); ge.load_package()
from graph_wavenet_b200 import engine as E, native as NV
from graph_wavenet_b200.metrics import StandardScaler
dev = torch.device("cuda:0")
lib = NV.get_lib()
PREC = {"tf32": NV.PREC_TF32, "fp32x3": NV.PREC_FP32X3}


def step_time(name, N, B, T, aptonly=False, blocks=4, layers=2, steps=10, tiers=("fp32x3", "tf32"), density=0.05, gflop=None):
    gen = torch.Generator().manual_seed(0)
    sup = None if aptonly else [s.to(dev) for s in O.synthetic_supports(N, density, gen)]
    x, y = O.synthetic_batch(B, N, T, 2, gen)
    x, y = x.to(dev), y.to(dev)
    for tier in tiers:
        torch.manual_seed(999)
        tr = E.trainer(StandardScaler(54.0, 20.0), 2, T, N, 32, 0.3, 1e-3, 1e-4, dev, sup, True, True, None, blocks, layers)
        tr.model.precision = PREC[tier]
        for _ in range(3):
            tr.train(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = tr.train(x, y)[0]
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        rec = {"config": name, "tier": tier, "N": N, "B": B, "T": T, "ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 1),
               "loss": round(loss, 4), "mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}
        if gflop:
            rec["model_TFLOPs"] = round(gflop / ms, 1)
        print(json.dumps(rec), flush=True)
        del tr
        torch.cuda.empty_cache()


def contraction(V, B, L, iters=5):
    gen = torch.Generator().manual_seed(V)
    ld = (V + 3) // 4 * 4
    S = torch.zeros(V, ld)
    S[:, :V] = torch.softmax(torch.randn(V, V, generator=gen), dim=1)
    x = torch.randn(B, L, V, 32, generator=gen).to(dev)
    Sd = S.to(dev)
    y = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        lib.check(lib.dll.gwn_node_contract(x.data_ptr(), Sd.data_ptr(), ld, y.data_ptr(), B, L, V, 32, NV.PREC_TF32, st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.check(lib.dll.gwn_node_contract(x.data_ptr(), Sd.data_ptr(), ld, y.data_ptr(), B, L, V, 32, NV.PREC_TF32, st))
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(json.dumps({"op": "node_contract tf32 (tcgen05)", "V": V, "B": B, "L": L, "us": round(ms * 1e3, 1),
                      "TFLOPs": round(2.0 * B * L * 32 * V * V / ms / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    if len(sys.argv) > 1 and sys.argv[1] == "contraction_only":
        contraction(2048, 8, 24, iters=2)
        sys.exit(0)
    step_time("METR-LA", 207, 64, 12, gflop=217.4)
    step_time("PEMS-BAY aptonly", 325, 64, 12, aptonly=True, gflop=249.9)
    step_time("CRASH N=200 T=12", 200, 64, 12, gflop=205.9)
    step_time("CRASH N=200 B=32 T=48", 200, 32, 48, gflop=834.1)
    contraction(2048, 8, 24)
    contraction(4096, 4, 24)
    if not quick:
        step_time("large N=2048 8x2", 2048, 64, 12, blocks=8, layers=2, steps=3, density=16 / 2048, gflop=51131.0)
