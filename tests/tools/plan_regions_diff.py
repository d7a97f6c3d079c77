"""GPU: dump every workspace / scratch region and the flat gradient of one forward + backward pass of the plan (fp32x3 tier,
METR-LA-like configuration) to a file, or compare a run with such a file region by region (same probe).  Used to localise
the addend-buffer race of tcpos.cuh: two builds / two runs must agree to ~1e-7 in every region.
usage: python tests/tools/plan_regions_diff.py N B out.pt            (save)
       python tests/tools/plan_regions_diff.py N B - ref.pt          (compare with ref.pt)"""
import ctypes, os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
from oracle import gwnet_oracle as O
ge.build(); ge.load_package()
from graph_wavenet_b200 import native as NV
from graph_wavenet_b200.runtime import PlanRunner, make_config
N, B, outp = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
gpu = NV.get_lib()
dev = torch.device("cuda:0")
cfg = O.GwnetConfig(num_nodes=N, dropout=0.0, n_static_supports=2, has_supports=True)
gen = torch.Generator().manual_seed(0)
sup = O.synthetic_supports(N, 0.05, gen)
x, _ = O.synthetic_batch(B, N, 12, 2, gen)
x = torch.nn.functional.pad(x, (1, 0, 0, 0))
torch.manual_seed(999)
st = O.init_state(cfg)
c = make_config(batch=B, num_nodes=N, seq_len=13, in_dim=2, out_dim=12, residual_channels=32, dilation_channels=32,
                skip_channels=256, end_channels=512, kernel_size=2, blocks=4, layers=2, n_static_supports=2,
                gcn_bool=1, adaptive=1, gcn=1, dropout=0.0, precision=NV.PREC_FP32X3)
r = PlanRunner(gpu, c)
params = [st[k].clone().contiguous().to(dev) for k in r.plan.names]
ws = torch.zeros(r.plan.fwd_bytes, dtype=torch.uint8, device=dev)
r._scratch = torch.zeros(r.plan.bwd_bytes, dtype=torch.uint8, device=dev)
out, ctx = r.forward(params, [s.to(dev) for s in sup], x.to(dev), True, workspace=ws)
probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(1)) * O.relu_safe_positions({k: v.detach().cpu().clone() for k, v in st.items()}, cfg, x, sup, True)
if len(sys.argv) > 4:
    probe = torch.load(sys.argv[4])['probe']
gflat, _ = r.backward(ctx, params, probe.to(dev))
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 16)
r.lib.dll.gwn_plan_debug_layout(r.plan.handle, buf, 1 << 16)
layout = buf.value.decode()
if len(sys.argv) > 4:
    ref = torch.load(sys.argv[4])
    wsf, scf = ws.view(torch.float32), r._scratch.view(torch.float32)
    for line in layout.strip().split("\n"):
        space, name, off, n = line.split(); off, n = int(off), int(n)
        a = (ref["ws"] if space == "fwd" else ref["sc"])[off:off+n].to(dev).double()
        b = (wsf if space == "fwd" else scf)[off:off+n].double()
        err = (a - b).norm().item() / max(a.norm().item(), 1e-30)
        bad = ((a - b).abs() > 1e-4 * (a.abs().max().item() + 1e-30)).nonzero().flatten()
        if err > 1e-6:
            print(f"{space} {name:8s} off={off} n={n} rel={err:.2e} nbad={bad.numel()} first={bad[:6].tolist()} last={bad[-2:].tolist()}")
    print("probe diff", (ref["probe"] - probe).abs().max().item(), "nonzero", int((ref["probe"] != probe).sum()))
    print("grad rel", ((ref["g"].to(dev).double() - gflat.double()).norm() / ref["g"].double().norm()).item())
else:
    torch.save({"ws": ws.view(torch.float32).cpu(), "sc": r._scratch.view(torch.float32).cpu(), "g": gflat.cpu(), "probe": probe}, outp)
    print("saved")
