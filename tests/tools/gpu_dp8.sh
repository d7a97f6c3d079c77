set -x
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 3 > gpurun_out/bench_dp8.log 2>&1; tail -c 1000 gpurun_out/bench_dp8.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 100 --warmup 3 > gpurun_out/bench_dp4.log 2>&1; tail -c 600 gpurun_out/bench_dp4.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_dp2.log 2>&1; tail -c 600 gpurun_out/bench_dp2.log
