set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 100 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_tf32.log 2>&1; tail -c 300 gpurun_out/bench_tf32.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_tf32.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-roofline --no-graph > gpurun_out/ncu_l.log 2>&1
# tcpos kernels of layer 0: forward gate (1st tcpos launch), mlp (2nd); backward kernels come later in the step
ncu --set full --clock-control none --import-source on -k regex:tcpos_kernel -s 32 -c 2 -o gpurun_out/full_tcpos_fwd -f python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-roofline --no-graph > gpurun_out/ncu_full_tcpos_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tcpos_kernel -s 69 -c 4 -o gpurun_out/full_tcpos_bwd -f python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-roofline --no-graph > gpurun_out/ncu_full_tcpos_bwd.log 2>&1
ls -la gpurun_out
