set -x
python tests/tools/tf32_rounding_probe.py 2>&1 | tail -4
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 100 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_tf32.log 2>&1; tail -c 300 gpurun_out/bench_tf32.log
