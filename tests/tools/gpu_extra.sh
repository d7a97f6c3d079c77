timeout 600 python tests/tools/configs_bench.py > gpurun_out/configs_bench_r01g.log 2>&1; grep -E "^\{" gpurun_out/configs_bench_r01g.log | tail -12 | cut -c1-200
GWNET_B200_PDL=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_tf32.csv python bench.py --precision tf32 --steps 2 --warmup 1 --skip-cpu-baseline --skip-roofline --skip-tiers --no-graph > gpurun_out/ncu_l2.log 2>&1
tail -c 200 gpurun_out/ncu_l2.log
