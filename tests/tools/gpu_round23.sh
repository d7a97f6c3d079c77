set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 100 --skip-cpu-baseline > gpurun_out/bench_default.log 2>&1; tail -c 200 gpurun_out/bench_default.log
