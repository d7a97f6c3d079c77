set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python tests/tools/tc_check.py > gpurun_out/tc_check.log 2>&1; tail -8 gpurun_out/tc_check.log
python tests/tools/configs_bench.py > gpurun_out/configs_bench.log 2>&1; grep -E "^\{" gpurun_out/configs_bench.log | tail -6
