"""2-GPU check (torchrun): after N data-parallel steps (NCCL all-reduce captured inside the CUDA graph) the replicas hold
bit-identical parameters, and they equal a single-GPU run that processes both shards itself and averages the gradients
is NOT expected (BatchNorm statistics are local) -- so the reference here is rank-0's own loss curve vs the loss of the
same shard trained alone for step 0 (identical, same init) and replica equality afterwards."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import __graft_entry__ as ge
from oracle import gwnet_oracle as O
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
print(f"[rank {rank}] init", flush=True)
dist.init_process_group("nccl", device_id=dev)
print(f"[rank {rank}] pg up", flush=True)
ge.build(); ge.load_package()
from graph_wavenet_b200 import engine as E
from graph_wavenet_b200.metrics import StandardScaler
gen = torch.Generator().manual_seed(0)
sup = [s.to(dev) for s in O.synthetic_supports(207, 0.05, gen)]
torch.manual_seed(999 + rank)                     # different init per rank: the broadcast must fix it
tr = E.trainer(StandardScaler(54.0, 20.0), 2, 12, 207, 32, 0.0, 1e-3, 1e-4, dev, sup, True, True, None)
tr.enable_data_parallel()
print(f"[rank {rank}] broadcast done", flush=True)
g2 = torch.Generator().manual_seed(100 + rank)
x, y = O.synthetic_batch(32, 207, 12, 2, g2)
x, y = x.to(dev), y.to(dev)
losses = []
for i in range(5):
    losses.append(tr.train(x, y)[0])
    print(f"[rank {rank}] step {i} loss {losses[-1]:.4f}", flush=True)
flat = tr.model._flat.param.clone()
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
same = all(torch.equal(gathered[0], g) for g in gathered)
st = next(iter(tr._steps.values()))
if rank == 0:
    print("graph captured:", st.graph is not None, "| losses:", [round(l, 4) for l in losses], "| replicas bit-identical after 5 steps:", same, flush=True)
assert same and st.graph is not None
dist.barrier(); dist.destroy_process_group()
