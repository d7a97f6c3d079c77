GWNET_B200_PDL=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r01g.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-roofline --skip-tiers --no-graph > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log | cut -c1-300
