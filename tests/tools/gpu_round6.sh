set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 3 --skip-cpu-baseline --precision fp32x3 > gpurun_out/bench_fp32x3.log 2>&1; tail -c 300 gpurun_out/bench_fp32x3.log
ncu --set full --clock-control none --import-source on -k regex:tcpos_kernel -s 39 -c 2 -o gpurun_out/full_tcpos_l0 -f python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-roofline --no-graph > gpurun_out/ncu_full_tcpos_l0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nconv_tc_kernel -s 37 -c 2 -o gpurun_out/full_nconv_l0 -f python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-roofline --no-graph > gpurun_out/ncu_full_nconv_l0.log 2>&1
