set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 3 --skip-cpu-baseline --precision fp32x3 > gpurun_out/bench_fp32x3.log 2>&1; tail -c 300 gpurun_out/bench_fp32x3.log
python bench.py --steps 50 --warmup 3 --skip-cpu-baseline --precision tf32 > gpurun_out/bench_tf32.log 2>&1; tail -c 300 gpurun_out/bench_tf32.log
