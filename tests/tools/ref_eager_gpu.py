"""GPU context number (not a bench value): the oracle port of the reference -- the same torch ops as
model.py / engine.py -- run by PyTorch eager ON the B200 (cuBLAS / cuDNN), i.e. the library path the
reference itself would take on this box.  Prints ms/step for trainer.train at the METR-LA shape."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
from oracle import gwnet_oracle as O

dev = torch.device("cuda:0")
for allow_tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = allow_tf32          # reference default: True
    torch.backends.cuda.matmul.allow_tf32 = False         # reference default
    cfg = O.GwnetConfig(num_nodes=207, dropout=0.3, n_static_supports=2)
    gen = torch.Generator().manual_seed(0)
    sup = [s.to(dev) for s in O.synthetic_supports(207, 0.05, gen)]
    torch.manual_seed(999)
    st = {k: v.to(dev) for k, v in O.init_state(cfg).items()}
    tr = O.OracleTrainer(cfg, st, sup, 54.0, 20.0)
    x, y = O.synthetic_batch(64, 207, 12, 2, gen)
    x, y = x.to(dev), y.to(dev)
    for _ in range(5):
        tr.train(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    n = 20
    for _ in range(n):
        tr.train(x, y)
    e1.record(); torch.cuda.synchronize()
    print(f"reference ops, torch eager on GPU, cudnn.allow_tf32={allow_tf32}: {e0.elapsed_time(e1)/n:.2f} ms/step "
          f"({64/(e0.elapsed_time(e1)/n*1e-3):.0f} samples/s), wall {(time.perf_counter()-t0)/n*1e3:.2f} ms/step", flush=True)
