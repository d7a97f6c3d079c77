# one step, every launch: duration, DRAM bytes, L2 bytes, tensor-pipe activity (a handful of metrics: few replays)
GWNET_B200_PDL=0 GWNET_B200_SIDE_STREAM=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum \
  --clock-control none -c 700 --csv --log-file gpurun_out/metrics_step.csv \
  python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-roofline --skip-tiers --no-graph > gpurun_out/ncu_m.log 2>&1
tail -c 300 gpurun_out/ncu_m.log
