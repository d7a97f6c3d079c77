set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tests/tools/configs_bench.py > gpurun_out/configs_bench.log 2>&1; grep -E "^\{|Error|error" gpurun_out/configs_bench.log | tail -20
