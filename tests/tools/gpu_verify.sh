set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 400 gpurun_out/bench_default.log
