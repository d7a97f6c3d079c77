"""CPU, world_size 2 over gloo: the data-parallel plumbing of ``engine.trainer`` -- rank-0
broadcast of the initial state, batch shards, ONE all-reduce over the flat gradient buffer, the 1/world scale
folded into the fused clip+Adam pass.  The GPU-less container cannot run the CUDA kernels, so the test injects
the host-emulation build of the same sources (tests/_hostemu, test infrastructure) as the library behind the
engine; what is under test is the host logic, which is identical on the GPU.

Expected values: the oracle run on each shard, gradients averaged, clip 5, torch.optim.Adam (the definition of
DP parity in SURVEY.md section 8(e): per-shard reference, averaged gradients, local BatchNorm statistics)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, emu_path, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    import __graft_entry__ as ge
    ge.load_package()
    from graph_wavenet_b200 import native as N, engine as E
    from graph_wavenet_b200.metrics import StandardScaler
    from helpers import load_case
    N._LIB = N.Lib(emu_path)                       # TEST-ONLY stand-in for the CUDA library
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rec = load_case("dbl_adp")
    cfg = rec["cfg"]
    dev = torch.device("cpu")
    torch.manual_seed(1000 + rank)                 # different init per rank: the broadcast must fix it
    tr = E.trainer(StandardScaler(54.0, 20.0), cfg.in_dim, cfg.out_dim, cfg.num_nodes, cfg.residual_channels, 0.0,
                   1e-3, 1e-4, dev, rec["supports"], cfg.gcn_bool, cfg.addaptadj, None, cfg.blocks, cfg.layers)
    if rank == 0:
        tr.model.load_state_dict(rec["state0"])
    tr.enable_data_parallel()
    x, y = rec["x"], rec["y"][:, :, : cfg.out_dim]
    half = x.shape[0] // world
    xs, ys = x[rank * half:(rank + 1) * half], y[rank * half:(rank + 1) * half]
    metrics = [tr.train(xs, ys) for _ in range(2)]
    state = {k: v.detach().clone() for k, v in tr.model.state_dict().items()}
    torch.save({"metrics": metrics, "state": state}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_step_matches_per_shard_reference(tmp_path):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    from helpers import load_case, assert_close_rel
    from oracle import gwnet_oracle as O
    emu_path = ge.build_hostemu(os.path.join(HERE, "_hostemu"))
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, emu_path, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]

    # expected: per-shard oracle, averaged gradients, clip 5, Adam(lr 1e-3, wd 1e-4); BatchNorm buffers stay local
    rec = load_case("dbl_adp")
    cfg = rec["cfg"]
    x, y = rec["x"], rec["y"][:, :, : cfg.out_dim]
    half = x.shape[0] // world
    pk = [k for k in rec["state0"] if not O.is_buffer(k)]
    params = {k: rec["state0"][k].clone().requires_grad_(True) for k in pk}
    bufs = [{k: v.clone() for k, v in rec["state0"].items() if O.is_buffer(k)} for _ in range(world)]
    opt = torch.optim.Adam([params[k] for k in pk], lr=1e-3, weight_decay=1e-4)
    want_metrics = [[], []]
    for step in range(2):
        grads = {k: None for k in pk}
        for r in range(world):
            st = dict(bufs[r])
            st.update(params)
            for k in pk:
                params[k].grad = None
            inp = torch.nn.functional.pad(x[r * half:(r + 1) * half], (1, 0, 0, 0))
            out = O.forward(st, cfg, inp, rec["supports"], True).transpose(1, 3)
            pred = out * 20.0 + 54.0
            real = y[r * half:(r + 1) * half].unsqueeze(1)
            loss = O.masked_mae(pred, real, 0.0)
            loss.backward()
            want_metrics[r].append((loss.item(), O.masked_mape(pred, real, 0.0).item(), O.masked_rmse(pred, real, 0.0).item()))
            for k in pk:
                if params[k].grad is not None:
                    grads[k] = params[k].grad.clone() if grads[k] is None else grads[k] + params[k].grad
            for k in bufs[r]:
                bufs[r][k] = st[k].detach().clone()
        for k in pk:
            params[k].grad = None if grads[k] is None else grads[k] / world
        torch.nn.utils.clip_grad_norm_([params[k] for k in pk], 5)
        opt.step()

    for r in range(world):
        for got, want in zip(res[r]["metrics"], want_metrics[r]):
            for a, b in zip(got, want):
                assert abs(a - b) <= 1e-4 * abs(b) + 1e-6, (r, got, want)
        for k in pk:
            assert_close_rel(res[r]["state"][k], params[k].detach(), 2e-3, f"rank {r} param {k}", floor=1e-5)
    for k in pk:      # replicas stay bit-identical: same averaged gradient, same update
        assert torch.equal(res[0]["state"][k], res[1]["state"][k]), k
