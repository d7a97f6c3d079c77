# final measurements of the round (small artefacts only; the ncu captures are in gpu_final_ncu.sh)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 300 gpurun_out/bench_default.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -c 300 gpurun_out/bench_reference.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
GWNET_B200_PDL=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_fp32x3.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-roofline --skip-tiers --no-graph > gpurun_out/ncu_l.log 2>&1
