set -x
export NCCL_DEBUG=WARN
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/tools/dp_check.py > gpurun_out/dp_check.log 2>&1; tail -25 gpurun_out/dp_check.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_dp2.log 2>&1; tail -c 1500 gpurun_out/bench_dp2.log
