"""N-GPU check (torchrun): data-parallel trainer.train on the fused CUDA-graph step, once with the peer-memory gradient
exchange (gwn_allreduce_adam_step: one kernel over NVLink inside the step's single graph) and once with the NCCL
all-reduce between two graphs.  For each mode: after 5 steps the replicas must hold bit-identical parameters (different
seeds per rank before the broadcast); the two modes must agree with each other to rounding (they sum the ranks in a
different order); and rank 0's first-step loss must equal the same shard trained alone (same init: BatchNorm statistics
are per shard, DDP semantics).  Prints one JSON line per mode with the timing of 50 further steps.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/tools/dp_check.py"""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import __graft_entry__ as ge
from oracle import gwnet_oracle as O
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
ge.build(); ge.load_package()
from graph_wavenet_b200 import engine as E, native as NV
from graph_wavenet_b200.metrics import StandardScaler
gen = torch.Generator().manual_seed(0)
sup = [s.to(dev) for s in O.synthetic_supports(207, 0.05, gen)]
g2 = torch.Generator().manual_seed(100 + rank)
x, y = O.synthetic_batch(64, 207, 12, 2, g2)
x, y = x.to(dev), y.to(dev)


def run(mode):
    os.environ["GWNET_B200_P2P_ALLREDUCE"] = "1" if mode == "p2p" else "0"
    torch.manual_seed(999 + rank)                     # different init per rank: the broadcast must fix it
    tr = E.trainer(StandardScaler(54.0, 20.0), 2, 12, 207, 32, 0.0, 1e-3, 1e-4, dev, sup, True, True, None)
    tr.enable_data_parallel()
    losses = [tr.train(x, y)[0] for _ in range(5)]
    st = next(iter(tr._steps.values()))
    flat = tr.model._flat.param.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    torch.cuda.synchronize(dev); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        tr.train(x, y)
    e1.record(); torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / 200], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    flag = NV.get_lib().dll.gwn_tc_error_flag(1)
    rec = {"mode": mode, "world": world, "p2p_active": st.p2p is not None, "graphs_per_step": 2 if st.graph_tail is not None else 1,
           "losses": [round(l, 5) for l in losses], "replicas_bit_identical": same, "ms_per_step": round(ms.item(), 4),
           "timeout_flag": flag}
    if rank == 0:
        print(json.dumps(rec), flush=True)
    assert same and st.graph is not None and flag == 0
    return losses, flat


lp, fp = run("p2p")
ln, fn = run("nccl")
run("p2p")          # timing again in the other order (clocks / caches warm for both)
run("nccl")
rel = float((fp - fn).norm() / fn.norm())
if rank == 0:
    print(json.dumps({"p2p_vs_nccl_param_rel_l2_after_5_steps": rel, "loss_diff": max(abs(a - b) for a, b in zip(lp, ln))}), flush=True)
assert rel < 1e-4
# the same shard alone, same (broadcast) init: step-0 loss is identical (the forward does not depend on the exchange)
if rank == 0:
    torch.manual_seed(999)
    tr1 = E.trainer(StandardScaler(54.0, 20.0), 2, 12, 207, 32, 0.0, 1e-3, 1e-4, dev, sup, True, True, None)
    l0 = tr1.train(x, y)[0]
    print(json.dumps({"single_gpu_step0_loss": round(l0, 5), "dp_step0_loss": lp[0]}), flush=True)
    assert abs(l0 - lp[0]) <= 1e-5 * abs(l0)
dist.barrier(); dist.destroy_process_group()
