"""Differential debugging: run the same plan on the GPU (CUDA kernels) and in the host emulation
(same functors as serial loops) and diff every workspace / scratch region.
usage: python tests/tools/diff_emu_gpu.py N B aptonly(0/1)"""
import ctypes, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
from oracle import gwnet_oracle as O

ge.build(); ge.load_package()
from graph_wavenet_b200 import native as NV
from graph_wavenet_b200.runtime import PlanRunner, make_config

N, B, apt = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
emu = NV.Lib(ge.build_hostemu(os.path.join(ROOT, "tests", "_hostemu")))
gpu = NV.get_lib()
dev = torch.device("cuda:0")
cfg = O.GwnetConfig(num_nodes=N, dropout=0.0, n_static_supports=0 if apt else 2, has_supports=not apt)
gen = torch.Generator().manual_seed(0)
sup = None if apt else O.synthetic_supports(N, 0.05, gen)
x, _ = O.synthetic_batch(B, N, 12, 2, gen)
x = torch.nn.functional.pad(x, (1, 0, 0, 0))
torch.manual_seed(999)
st = O.init_state(cfg)
c = make_config(batch=B, num_nodes=N, seq_len=13, in_dim=2, out_dim=12, residual_channels=32, dilation_channels=32,
                skip_channels=256, end_channels=512, kernel_size=2, blocks=4, layers=2, n_static_supports=0 if apt else 2,
                gcn_bool=1, adaptive=1, gcn=1, dropout=0.0)
res = {}
for name, lib, d in (("emu", emu, torch.device("cpu")), ("gpu", gpu, dev)):
    r = PlanRunner(lib, c)
    params = [st[k].clone().contiguous().to(d) for k in r.plan.names]
    ws = torch.zeros(r.plan.fwd_bytes, dtype=torch.uint8, device=d)
    r._scratch = torch.zeros(r.plan.bwd_bytes, dtype=torch.uint8, device=d)
    out, ctx = r.forward(params, [s.to(d) for s in sup] if sup else None, x.to(d), True, workspace=ws)
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    gflat, _ = r.backward(ctx, params, probe.to(d))
    if d.type == "cuda":
        torch.cuda.synchronize()
    res[name] = (ws.cpu().view(torch.float32), r._scratch.cpu().view(torch.float32), gflat.cpu(), r)
buf = ctypes.create_string_buffer(1 << 16)
r = res["gpu"][3]
r.lib.dll.gwn_plan_debug_layout(r.plan.handle, buf, 1 << 16)
for line in buf.value.decode().strip().split("\n"):
    space, name, off, n = line.split()
    off, n = int(off), int(n)
    a = res["emu"][0 if space == "fwd" else 1][off:off + n].double()
    b = res["gpu"][0 if space == "fwd" else 1][off:off + n].double()
    err = (a - b).norm().item() / max(a.norm().item(), 1e-30)
    bad = ((a - b).abs() > 1e-4 * (a.abs().max().item() + 1e-30)).nonzero().flatten()
    flag = "  <<<<" if err > 1e-4 else ""
    extra = f" nbad={bad.numel()} first={bad[:4].tolist()} last={bad[-2:].tolist()}" if bad.numel() else ""
    print(f"{space} {name:8s} off={off:10d} n={n:10d} rel={err:.2e}{extra}{flag}")
ge_, gg = res["emu"][2].double(), res["gpu"][2].double()
print("grad_flat rel", ((ge_ - gg).norm() / ge_.norm()).item())
