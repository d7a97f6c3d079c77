"""GPU: fixed cost of one node-contraction launch -- time gwn_node_contract_x3 / gwn_node_contract (tf32) at the METR-LA
graph size over a sweep of slab counts, back to back on one stream (warm L2), and fit time = a + b * rounds.
usage: python tests/tools/nconv_sweep.py [V]"""
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
V = int(sys.argv[1]) if len(sys.argv) > 1 else 207
ld = (V + 3) // 4 * 4
S = torch.zeros(V, ld)
S[:, :V] = torch.softmax(torch.randn(V, V), dim=1)
S = S.to(dev)
Slo = torch.empty_like(S)
lib.check(lib.dll.gwn_split_lo(S.data_ptr(), Slo.data_ptr(), S.numel(), st))
for tier in ("fp32x3", "tf32"):
    for L in (1, 2, 3, 4, 6, 9, 12, 18, 24, 37, 48, 74, 96):
        B = 64
        x = torch.randn(B, L, V, 32, device=dev)
        y = torch.empty_like(x)
        def run():
            if tier == "tf32":
                lib.check(lib.dll.gwn_node_contract(x.data_ptr(), S.data_ptr(), ld, y.data_ptr(), B, L, V, 32, NV.PREC_TF32, st))
            else:
                lib.check(lib.dll.gwn_node_contract_x3(x.data_ptr(), S.data_ptr(), Slo.data_ptr(), ld, y.data_ptr(), B, L, V, 32, st))
        for _ in range(5):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 50
        e0.record()
        for _ in range(iters):
            run()
        e1.record(); e1.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        tiles = (B * L + 7) // 8
        print(json.dumps({"tier": tier, "V": V, "slabs": B * L, "row_tiles": tiles, "rounds_74": round(tiles / 74, 2), "us": round(us, 2),
                          "flag": lib.dll.gwn_tc_error_flag(1)}), flush=True)
