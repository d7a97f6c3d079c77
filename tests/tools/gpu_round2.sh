set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_tf32.log 2>&1; tail -c 1500 gpurun_out/bench_tf32.log
python bench.py --steps 30 --warmup 3 --no-graph --skip-cpu-baseline --skip-roofline > gpurun_out/bench_tf32_nograph.log 2>&1; tail -c 600 gpurun_out/bench_tf32_nograph.log
